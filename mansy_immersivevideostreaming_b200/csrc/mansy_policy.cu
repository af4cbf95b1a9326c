// mansy_policy.cu -- policy / value forward that consumes the simulator's observation rows.
//
// Reference: bitrate_selection/models/mansy.py:26-51 (FeatureNet), :63-66 (Actor), :77-80 (Critic);
// bitrate_selection/models/simple_rl.py:21-35,46-49,60-63; Categorical sampling
// bitrate_selection/run_mansy.py:228-229.
//
// Every Conv1d in the reference has kernel_size == input length, i.e. it is a Linear over the
// flattened (channels x length) segment of the observation row, so the whole network is
//   layer 1: 10 (MANSY) / 5 (SimpleRL) independent Linear(K_b -> 128) + LeakyReLU on row segments,
//   layer 2: Linear(F -> 128) + LeakyReLU for the actor and for the critic (stacked: F -> 256),
//            MANSY adds the qoe_weight branch features as a residual,
//   heads  : Linear(128 -> 15) and Linear(128 -> 1); SimpleRL's actor returns softmax probabilities.
// The FeatureNet instance is shared by actor and critic (run_mansy.py:207-209), so its output is
// computed once.
//
// This file holds the FP32 CUDA-core implementation (exact fp32 FMA math, used for parity against
// a torch fp32 reference).  Weights are re-laid out K-major at create time so that consecutive
// threads read consecutive weights.
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mansy_policy.cuh"

namespace mansy {

constexpr int kEnvTile = 16;          // environments per CTA
constexpr int kPolicyThreads = 256;

// One CTA = 16 environments.  Dynamic shared memory:
//   feat [feat_dim][16]  layer-1 activations
//   xs   [320][16]       staged observation segment of the current branch
//   hid  [256][16]       layer-2 activations (actor | critic)
__global__ void __launch_bounds__(kPolicyThreads)
policy_forward_kernel(const __grid_constant__ PolicyDev P, const float *__restrict__ obs, int64_t obs_stride, int n,
                      float *__restrict__ logits, float *__restrict__ value) {
  extern __shared__ __align__(16) float smem[];
  float *feat = smem;
  float *xs = feat + (size_t)P.feat_dim * kEnvTile;
  float *hid = xs + 320 * kEnvTile;
  const int t = threadIdx.x;
  const int env0 = blockIdx.x * kEnvTile;
  const int n_here = min(kEnvTile, n - env0);

  // ---- layer 1 -------------------------------------------------------------------------
  const int f = t & (kHidden - 1), half = t >> 7;        // 128 features x 2 halves of 8 envs
  for (int b = 0; b < P.n_branches; ++b) {
    const int K = P.k[b];
    for (int idx = t; idx < K * kEnvTile; idx += kPolicyThreads) {
      const int e = idx / K, kk = idx - e * K;
      xs[kk * kEnvTile + e] = e < n_here ? __ldg(obs + (size_t)(env0 + e) * obs_stride + P.obs_off[b] + kk) : 0.f;
    }
    __syncthreads();
    float acc[8];
    const float bias = __ldg(P.b1 + b * kHidden + f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias;
    const float *w = P.w1t + P.w_off[b] + f;
    for (int kk = 0; kk < K; ++kk) {
      const float wv = __ldg(w + (size_t)kk * kHidden);
      const float4 x0 = *reinterpret_cast<const float4 *>(xs + kk * kEnvTile + half * 8);
      const float4 x1 = *reinterpret_cast<const float4 *>(xs + kk * kEnvTile + half * 8 + 4);
      acc[0] = fmaf(wv, x0.x, acc[0]); acc[1] = fmaf(wv, x0.y, acc[1]);
      acc[2] = fmaf(wv, x0.z, acc[2]); acc[3] = fmaf(wv, x0.w, acc[3]);
      acc[4] = fmaf(wv, x1.x, acc[4]); acc[5] = fmaf(wv, x1.y, acc[5]);
      acc[6] = fmaf(wv, x1.z, acc[6]); acc[7] = fmaf(wv, x1.w, acc[7]);
    }
    float *dst = feat + (size_t)(b * kHidden + f) * kEnvTile + half * 8;
    *reinterpret_cast<float4 *>(dst) = make_float4(leaky(acc[0]), leaky(acc[1]), leaky(acc[2]), leaky(acc[3]));
    *reinterpret_cast<float4 *>(dst + 4) = make_float4(leaky(acc[4]), leaky(acc[5]), leaky(acc[6]), leaky(acc[7]));
    __syncthreads();
  }

  // ---- layer 2: thread t owns hidden unit t (0..127 actor, 128..255 critic) for 16 envs ------
  {
    float acc[kEnvTile];
    const float bias = __ldg(P.bfc + t);
#pragma unroll
    for (int i = 0; i < kEnvTile; ++i) acc[i] = bias;
    const float *w = P.wfct + t;
    for (int kk = 0; kk < P.feat_dim; ++kk) {
      const float wv = __ldg(w + (size_t)kk * 256);
      const float4 *x = reinterpret_cast<const float4 *>(feat + (size_t)kk * kEnvTile);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = x[j];
        acc[4 * j + 0] = fmaf(wv, v.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(wv, v.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(wv, v.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(wv, v.w, acc[4 * j + 3]);
      }
    }
    const float *res = P.residual_branch >= 0
                           ? feat + (size_t)(P.residual_branch * kHidden + (t & (kHidden - 1))) * kEnvTile
                           : nullptr;
#pragma unroll
    for (int i = 0; i < kEnvTile; ++i) {
      float hv = leaky(acc[i]);
      if (res) hv += res[i];                       // mansy.py:65,79: fc(features) + qoe_features
      hid[t * kEnvTile + i] = hv;
    }
  }
  __syncthreads();

  // ---- heads: thread (e, o): o < 15 actor logit, o == 15 critic value ---------------------
  {
    const int e = t & (kEnvTile - 1), o = t >> 4;
    const float *w = P.wout + o * kHidden;
    const float *hsrc = hid + (o == 15 ? kHidden * kEnvTile : 0) + e;
    float acc = __ldg(P.bout + o);
    for (int j = 0; j < kHidden; ++j) acc = fmaf(__ldg(w + j), hsrc[j * kEnvTile], acc);
    xs[e * 16 + o] = acc;                          // reuse xs as [16 envs][16 outputs]
  }
  __syncthreads();
  if (t < kEnvTile && t < n_here) {
    const int e = t;
    float out[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) out[o] = xs[e * 16 + o];
    if (P.softmax == 2) {                          // QoE identifier: torch.sigmoid on its 3 outputs (mansy.py:141)
#pragma unroll
      for (int o = 0; o < 3; ++o) out[o] = 1.0f / (1.0f + expf(-out[o]));
    } else if (P.softmax) {                        // simple_rl.py:48
      float m = out[0];
#pragma unroll
      for (int o = 1; o < kActions; ++o) m = fmaxf(m, out[o]);
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < kActions; ++o) { out[o] = expf(out[o] - m); s += out[o]; }
#pragma unroll
      for (int o = 0; o < kActions; ++o) out[o] = out[o] / s;
    }
    if (value) value[env0 + e] = out[15];
    if (logits) {
      float4 *dst = reinterpret_cast<float4 *>(logits + (size_t)(env0 + e) * 16);
      dst[0] = make_float4(out[0], out[1], out[2], out[3]);
      dst[1] = make_float4(out[4], out[5], out[6], out[7]);
      dst[2] = make_float4(out[8], out[9], out[10], out[11]);
      dst[3] = make_float4(out[12], out[13], out[14], 0.f);
    }
  }
}

// Categorical(logits).sample(): see categorical_sample (mansy_policy.cuh).
__global__ void policy_sample_kernel(const float *__restrict__ logits, int n, int is_probs, uint64_t seed,
                                     int64_t step, int env_offset, int32_t *__restrict__ actions,
                                     float *__restrict__ logp) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float p[kActions];
  const float4 *src = reinterpret_cast<const float4 *>(logits + (size_t)e * 16);
  const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
  p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
  p[8] = c.x; p[9] = c.y; p[10] = c.z; p[11] = c.w; p[12] = d.x; p[13] = d.y; p[14] = d.z;
  int act;
  float lp;
  categorical_sample(p, is_probs, seed, (uint64_t)(env_offset + e), (uint64_t)step, act, lp);
  actions[e] = act;
  if (logp) logp[e] = lp;
}

// One CTA per table row (video, chunk): thread t computes feature t of the memoised branches (t >> 7 selects the
// branch: 0 = next-chunk sizes, 1 = next-chunk qualities) from the row of size_norm / qual_norm -- exactly what the
// observation carries at those columns (mansy_step.cuh emit_obs_tables) -- then hidden unit t of actor.fc | critic.fc.
__global__ void __launch_bounds__(256) policy_memo_kernel(const PolicyDev P, const float *__restrict__ size_norm,
                                                          const float *__restrict__ qual_norm, int n_memo, int b0, int b1,
                                                          float *__restrict__ memo) {
  __shared__ float x[2 * kTableRow];
  __shared__ float feat[2 * kHidden];
  const int row = blockIdx.x, t = threadIdx.x;
  for (int i = t; i < n_memo * kTableRow; i += 256)
    x[i] = i < kTableRow ? __ldg(size_norm + (size_t)row * kTableRow + i) : __ldg(qual_norm + (size_t)row * kTableRow + i - kTableRow);
  __syncthreads();
  const int which = t >> 7, f = t & (kHidden - 1);
  if (which < n_memo) {
    const int b = which == 0 ? b0 : b1;
    float acc = __ldg(P.b1 + b * kHidden + f);
    const float *w = P.w1t + P.w_off[b] + f;
    const float *xb = x + which * kTableRow;
    for (int k = 0; k < kTableRow; ++k) acc = fmaf(__ldg(w + (size_t)k * kHidden), xb[k], acc);
    feat[t] = leaky(acc);
  }
  __syncthreads();
  float acc = 0.f;
  for (int m = 0; m < n_memo; ++m) {
    const int b = m == 0 ? b0 : b1;
    const float *w = P.wfct + (size_t)b * kHidden * 256 + t;
    for (int k = 0; k < kHidden; ++k) acc = fmaf(__ldg(w + (size_t)k * 256), feat[m * kHidden + k], acc);
  }
  memo[(size_t)row * 256 + t] = acc;
}

int policy_memo_for(mansy_policy *p, const SimDev &S, void *stream, const float **memo_out) {
  *memo_out = nullptr;
  int n_memo = 0, b[2] = {0, 0};
  if (p->dev.kind == MANSY_OBS_MANSY && S.obs_mode == MANSY_OBS_MANSY) { n_memo = 2; b[0] = 1; b[1] = 2; }        // conv1d2, conv1d3
  else if (p->dev.kind == MANSY_OBS_SIMPLE && S.obs_mode == MANSY_OBS_SIMPLE) { n_memo = 1; b[0] = 1; }           // conv1d_2
  else return MANSY_OK;
  const int rows = S.n_videos * S.n_chunks;
  if (p->memo_key != S.uid || p->memo_rows != rows) {
    if (p->memo_cap < rows) {
      if (p->memo) { cudaDeviceSynchronize(); cudaFree(p->memo); p->memo = nullptr; p->memo_cap = 0; }
      void *d = nullptr;
      if (cudaMalloc(&d, (size_t)rows * 256 * sizeof(float)) != cudaSuccess) return set_error(MANSY_E_NOMEM, "cudaMalloc failed (policy memo)");
      p->memo = static_cast<float *>(d);
      p->memo_cap = rows;
    }
    p->memo_rows = rows;
    policy_memo_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(p->dev, S.size_norm, S.qual_norm, n_memo, b[0], b[1], p->memo);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_error(MANSY_E_CUDA, "policy_memo_kernel launch failed");
    p->memo_key = S.uid;
  }
  *memo_out = p->memo;
  return MANSY_OK;
}

}  // namespace mansy

using namespace mansy;

namespace {
int upload_f(mansy_policy *p, const std::vector<float> &host, const float **out) {
  void *d = nullptr;
  if (cudaMalloc(&d, host.size() * sizeof(float) + 16) != cudaSuccess) return set_error(MANSY_E_NOMEM, "cudaMalloc failed (policy)");
  p->allocs.push_back(d);
  if (cudaMemcpy(d, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    return set_error(MANSY_E_CUDA, "cudaMemcpy failed (policy)");
  *out = static_cast<const float *>(d);
  return MANSY_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------
// PPO learner inputs computed where the rollout lives
// ------------------------------------------------------------------------------------------
// utils/mansy_utils.py:42-49 + models/mansy_ppo.py:40-49.  F.mse_loss on 3 float32 values: squared differences
// summed in order, divided by 3; the blend is the float64 numpy expression of PPOPolicy.update.
__global__ void identifier_reward_kernel(const float *__restrict__ pred, const float *__restrict__ obs, int64_t obs_stride,
                                         const float *__restrict__ qoe_reward, double lamb, int n, float *__restrict__ ident,
                                         double *__restrict__ mixed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *w = obs + (size_t)i * obs_stride + 776;      // qoe_weight segment of the MANSY row (config.py)
  const float *p = pred + (size_t)i * 16;
  const float d0 = __fsub_rn(p[0], w[0]), d1 = __fsub_rn(p[1], w[1]), d2 = __fsub_rn(p[2], w[2]);
  const float mse = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), 3.0f);
  const float r = __fsub_rn(1.0f, mse);
  if (ident) ident[i] = r;
  if (mixed) mixed[i] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, lamb), (double)qoe_reward[i]), __dmul_rn(lamb, (double)r));
}

// tianshou 0.4.8 BasePolicy.compute_episodic_return / _gae_return (not vendored; restated in oracle/gae_oracle.py):
// one thread per environment walks its T transitions backwards; rows are [T][n] so every step is a coalesced read.
__global__ void gae_kernel(const float *__restrict__ reward, const float *__restrict__ value, const uint8_t *__restrict__ done,
                           const float *__restrict__ last_value, int T, int n, double gamma, double lam,
                           float *__restrict__ adv, float *__restrict__ ret) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double gae = 0.0;
  double v_next = (double)last_value[e];
  const double gl = __dmul_rn(gamma, lam);
  for (int t = T - 1; t >= 0; --t) {
    const size_t idx = (size_t)t * n + e;
    const double v = (double)value[idx];
    const double live = done[idx] ? 0.0 : 1.0;            // value_mask / end_flag
    const double delta = __dsub_rn(__dadd_rn((double)reward[idx], __dmul_rn(__dmul_rn(v_next, live), gamma)), v);
    gae = __dadd_rn(delta, __dmul_rn(__dmul_rn(live, gl), gae));
    if (adv) adv[idx] = (float)gae;
    if (ret) ret[idx] = (float)__dadd_rn(gae, v);
    v_next = v;
  }
}

namespace mansy {
int policy_device_of(mansy_policy_t p) { return p->device; }
}  // namespace mansy

extern "C" {

int mansy_policy_create(const mansy_policy_weights_t *w, int device, mansy_policy_t *out) {
  if (!out) return set_error(MANSY_E_INVALID, "out is NULL");
  *out = nullptr;
  if (!w) return set_error(MANSY_E_INVALID, "weights is NULL");
  if (w->kind != MANSY_OBS_MANSY && w->kind != MANSY_OBS_SIMPLE && w->kind != MANSY_NET_IDENTIFIER)
    return set_error(MANSY_E_INVALID, "bad policy kind");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return set_error(MANSY_E_CUDA, "no CUDA device available: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(MANSY_E_INVALID, "bad device index");
  DeviceScope dscope(device);
  if (dscope.err != cudaSuccess) return set_error(MANSY_E_CUDA, "cudaSetDevice failed");

  // branch -> (observation offset, K) in FeatureNet order (config.py row layouts)
  static const int mansy_off[10] = {0, 8, 328, 648, 728, 736, 744, 752, 779, 776};
  static const int mansy_k[10] = {8, 320, 320, 64, 8, 8, 8, 8, 1, 3};
  static const int simple_off[5] = {0, 8, 394, 392, 328};
  static const int simple_k[5] = {8, 320, 1, 2, 64};
  // QoE identifier (models/mansy.py:92-101): the same rows, fc2 reads the 15 action_one_hot floats instead of qoe_weight
  static const int ident_off[10] = {0, 8, 328, 648, 728, 736, 744, 752, 779, 760};
  static const int ident_k[10] = {8, 320, 320, 64, 8, 8, 8, 8, 1, 15};
  const bool is_ident = w->kind == MANSY_NET_IDENTIFIER;
  const bool is_mansy = w->kind == MANSY_OBS_MANSY || is_ident;      // MANSY observation rows
  const int nb = is_mansy ? 10 : 5;
  const int n_out = is_ident ? 3 : kActions;
  const int *off = is_ident ? ident_off : (is_mansy ? mansy_off : simple_off);
  const int *kk = is_ident ? ident_k : (is_mansy ? mansy_k : simple_k);
  for (int b = 0; b < nb; ++b)
    if (!w->branch_w[b] || !w->branch_b[b]) return set_error(MANSY_E_INVALID, "a branch weight pointer is NULL");
  if (!w->actor_fc_w || !w->actor_fc_b || !w->actor_out_w || !w->actor_out_b ||
      (!is_ident && (!w->critic_fc_w || !w->critic_fc_b || !w->critic_out_w || !w->critic_out_b)))
    return set_error(MANSY_E_INVALID, "a head weight pointer is NULL");

  mansy_policy *p = new (std::nothrow) mansy_policy();
  if (!p) return set_error(MANSY_E_NOMEM, "out of host memory");
  p->device = device;
  PolicyDev &d = p->dev;
  memset(&d, 0, sizeof(d));
  d.kind = w->kind; d.n_branches = nb; d.feat_dim = nb * kHidden;
  d.residual_branch = is_mansy ? 9 : -1;
  d.softmax = is_ident ? 2 : (is_mansy ? 0 : 1);
  const int F = d.feat_dim;

  std::vector<float> w1t, b1((size_t)nb * kHidden), wfct((size_t)F * 256), bfc(256), wout(16 * kHidden), bout(16);
  int woff = 0;
  for (int b = 0; b < nb; ++b) {
    d.obs_off[b] = off[b]; d.k[b] = kk[b]; d.w_off[b] = woff;
    w1t.resize((size_t)woff + (size_t)kk[b] * kHidden);
    for (int f = 0; f < kHidden; ++f) {
      for (int k = 0; k < kk[b]; ++k) w1t[(size_t)woff + (size_t)k * kHidden + f] = w->branch_w[b][(size_t)f * kk[b] + k];
      b1[(size_t)b * kHidden + f] = w->branch_b[b][f];
    }
    woff += kk[b] * kHidden;
  }
  for (int j = 0; j < kHidden; ++j) {
    for (int k = 0; k < F; ++k) {
      wfct[(size_t)k * 256 + j] = w->actor_fc_w[(size_t)j * F + k];
      wfct[(size_t)k * 256 + kHidden + j] = is_ident ? 0.f : w->critic_fc_w[(size_t)j * F + k];
    }
    bfc[j] = w->actor_fc_b[j];
    bfc[kHidden + j] = is_ident ? 0.f : w->critic_fc_b[j];
  }
  for (int o = 0; o < n_out; ++o) {
    for (int j = 0; j < kHidden; ++j) wout[(size_t)o * kHidden + j] = w->actor_out_w[(size_t)o * kHidden + j];
    bout[o] = w->actor_out_b[o];
  }
  if (!is_ident) {
    for (int j = 0; j < kHidden; ++j) wout[(size_t)15 * kHidden + j] = w->critic_out_w[j];
    bout[15] = w->critic_out_b[0];
  }

  int rc;
  if ((rc = upload_f(p, w1t, &d.w1t)) || (rc = upload_f(p, b1, &d.b1)) || (rc = upload_f(p, wfct, &d.wfct)) ||
      (rc = upload_f(p, bfc, &d.bfc)) || (rc = upload_f(p, wout, &d.wout)) || (rc = upload_f(p, bout, &d.bout))) {
    mansy_policy_destroy(p);
    return rc;
  }
  p->smem_bytes = ((size_t)F * kEnvTile + 320 * kEnvTile + 256 * kEnvTile) * sizeof(float);
  if (cudaFuncSetAttribute(policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes) !=
      cudaSuccess) {
    mansy_policy_destroy(p);
    return set_error(MANSY_E_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
  }
  tc_create(p, w);    // tensor-core images; on failure p->tc stays NULL and mansy_policy_forward_tc reports it
  *out = p;
  return MANSY_OK;
}

int mansy_policy_destroy(mansy_policy_t p) {
  if (!p) return MANSY_OK;
  DeviceScope dscope(p->device);
  tc_destroy(p);
  if (p->memo) cudaFree(p->memo);
  for (void *q : p->allocs) cudaFree(q);
  delete p;
  return MANSY_OK;
}

int mansy_policy_forward(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                         float *value_dev, void *stream) {
  if (!p || !obs_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (n < 0) return set_error(MANSY_E_INVALID, "n < 0");
  DeviceScope dscope(p->device);
  if (dscope.err != cudaSuccess) return set_error(MANSY_E_CUDA, "cudaSetDevice failed");
  const int need = p->dev.kind == MANSY_OBS_SIMPLE ? MANSY_OBS_SIMPLE_STRIDE : MANSY_OBS_MANSY_STRIDE;
  if (obs_stride < need - 4) return set_error(MANSY_E_INVALID, "obs_stride smaller than the observation row");
  if (logits_dev && (reinterpret_cast<uintptr_t>(logits_dev) & 15)) return set_error(MANSY_E_INVALID, "logits must be 16-byte aligned");
  if (n == 0) return MANSY_OK;
  const int grid = (n + kEnvTile - 1) / kEnvTile;
  policy_forward_kernel<<<grid, kPolicyThreads, p->smem_bytes, static_cast<cudaStream_t>(stream)>>>(
      p->dev, obs_dev, obs_stride, n, logits_dev, value_dev);
  count_launch();
  if (cudaGetLastError() != cudaSuccess) return set_error(MANSY_E_CUDA, "policy_forward_kernel launch failed");
  return MANSY_OK;
}

int mansy_policy_sample(const float *logits_dev, int32_t n, int32_t is_probs, uint64_t seed, int64_t step,
                        int32_t env_offset, int32_t *actions_dev, float *logp_dev, void *stream) {
  if (!logits_dev || !actions_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (n < 0) return set_error(MANSY_E_INVALID, "n < 0");
  if (n == 0) return MANSY_OK;
  DeviceScope dscope(device_of_pointer(logits_dev));
  policy_sample_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      logits_dev, n, is_probs, seed, step, env_offset, actions_dev, logp_dev);
  count_launch();
  if (cudaGetLastError() != cudaSuccess) return set_error(MANSY_E_CUDA, "policy_sample_kernel launch failed");
  return MANSY_OK;
}

int mansy_identifier_reward(const float *pred_dev, const float *obs_dev, int64_t obs_stride, const float *qoe_reward_dev,
                            double lamb, int32_t n, float *ident_dev, double *mixed_dev, void *stream) {
  if (!pred_dev || !obs_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (mixed_dev && !qoe_reward_dev) return set_error(MANSY_E_INVALID, "the blended reward needs qoe_reward");
  if (n < 0 || obs_stride < MANSY_OBS_MANSY_STRIDE) return set_error(MANSY_E_INVALID, "bad n / obs_stride");
  if (n == 0) return MANSY_OK;
  DeviceScope dscope(device_of_pointer(pred_dev));
  identifier_reward_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred_dev, obs_dev, obs_stride,
                                                                                          qoe_reward_dev, lamb, n, ident_dev,
                                                                                          mixed_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("identifier_reward_kernel: ") + cudaGetErrorString(e));
  return MANSY_OK;
}

int mansy_gae(const float *reward_dev, const float *value_dev, const uint8_t *done_dev, const float *last_value_dev,
              int32_t n_steps, int32_t n, double gamma, double lam, float *adv_dev, float *ret_dev, void *stream) {
  if (!reward_dev || !value_dev || !done_dev || !last_value_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (n_steps < 0 || n < 0) return set_error(MANSY_E_INVALID, "bad n_steps / n");
  if (n_steps == 0 || n == 0) return MANSY_OK;
  DeviceScope dscope(device_of_pointer(reward_dev));
  gae_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(reward_dev, value_dev, done_dev, last_value_dev,
                                                                            n_steps, n, gamma, lam, adv_dev, ret_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("gae_kernel: ") + cudaGetErrorString(e));
  return MANSY_OK;
}

}  // extern "C"
