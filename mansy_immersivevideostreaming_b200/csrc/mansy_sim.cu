// mansy_sim.cu -- sm_100a kernels and C ABI of the tile-based streaming simulator.
//
// Thread mapping: 8 lanes per environment (one lane per row of the 8x8 tile grid), 4 environments
// per warp, 16 per 128-thread CTA.  The scalar simulator chain (trace walk, buffer, QoE) is
// computed redundantly by the 8 lanes of a group -- it is a few hundred instructions against a
// 3.5 KB observation row -- while every table read and every observation write is cooperative:
// each warp-level 128-bit load/store touches four fully used 128-byte lines.  The kernel is an
// HBM-bound gather/scatter; there is nothing GEMM-shaped in it.
//
// Reference map (file:line relative to the reference root):
//   step_env / reset_episode / emit_obs: see mansy_step.cuh
//   viewport kernel viewport_prediction/utils/common.py:37-58,83-127, predict.py:33-48
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mansy_sim.cuh"

#ifndef MANSY_STEP_MIN_BLOCKS
#define MANSY_STEP_MIN_BLOCKS 5   // resident CTAs per SM the step kernel is compiled for (register cap 65536 / (128 * this)); 4 / 5 / 6 / 8 measured: 58.7 / 65.2 / 63.7 / 58.4 % of HBM peak at 1 Mi envs
#endif

namespace mansy {

// ------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

#define MANSY_CUDA(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return mansy::set_error(MANSY_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace mansy

#include "mansy_step.cuh"  // StepArgs, step_env, emit_obs, reset_episode, finish_episode (device code)

namespace mansy {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// One step of one environment group inside step_kernel.  `mask` is the shuffle mask of the group's collectives:
// the constant 0xFFFFFFFF when the whole warp is known to execute this together (plain SHFL), else the 8-lane
// group mask (partial-mask collectives cost ~10x the instructions).
template <int MODE>
__device__ __forceinline__ void step_once(const SimDev &S, const StepArgs &A, EnvState &st, float (&slot)[8], int e, int i,
                                          int sub, unsigned mask, int t) {
  const size_t r = (size_t)t * A.rows_per_step + i;
  float reward_f = 0.f;
  bool over = true;
  if (!(st.flags & kFlagFinished)) {
    const int action = A.action_mode == 0
                           ? __ldg(A.actions + i)
                           : hashed_action(A.seed, (uint64_t)(S.env_offset + e), (uint64_t)(A.step0 + t), kActions);
    const int slot_before = st.ep_step & 7;
    const double reward = step_env(S, st, slot, sub, mask, action, over,
                                   A.out.aux ? A.out.aux + r * MANSY_AUX_DOUBLES : nullptr,
                                   A.out.tile_versions ? A.out.tile_versions + r * kTiles : nullptr);
    reward_f = (float)reward;
    if (A.n_steps == 1 && sub == slot_before) store_slot(S, e, sub, slot);
    if (over) {
      finish_episode(S, e, st, sub);
      if (A.auto_reset) reset_episode(S, st);
      else st.flags |= kFlagFinished;
    }
  }
  if (sub == 0) {
    if (A.out.reward) A.out.reward[r] = reward_f;
    if (A.out.done) A.out.done[r] = over ? 1 : 0;
  }
  if (MODE != MANSY_OBS_NONE && A.out.obs) emit_obs<MODE>(S, st, slot, sub, mask, A.out.obs + r * A.out.obs_stride);
}

// (Measured and rejected, profiles/r02e_step_sweep.txt: the 2 560 / 1 280 table bytes of the row as bulk copies table ->
// shared memory -> row, 40 KB of staging per CTA: 79.4 % / 54.7 % of the HBM peak at 1 Mi envs against 82.7 % / 58.8 % for
// the cooperative 128-bit loads and stores below.)
template <int MODE>
__global__ void __launch_bounds__(kThreadsPerBlock, MANSY_STEP_MIN_BLOCKS)
step_kernel(const __grid_constant__ SimDev S, const __grid_constant__ StepArgs A) {
  // Programmatic dependent launch (no-ops without the launch attribute): the next kernel of the stream may be
  // scheduled right away -- it orders itself behind our completion -- and everything we read below (actions,
  // state, history) may have been written by the kernels before us.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int i = blockIdx.x * kEnvsPerBlock + (threadIdx.x >> 3);   // output row
  bool live = i < A.n;                                             // lanes past the end stay for the warp votes below
  const int sub = threadIdx.x & 7;
  const unsigned gmask = group_mask();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int e = live ? (A.env_ids ? __ldg(A.env_ids + i) : i) : 0;
  if (e < 0 || e >= S.n_envs) {            // the host wrappers validate ids (the reference raises IndexError); never index out of range
    if (sub == 0) atomicExch(S.error_flag, 2);
    live = false;
    e = 0;
  }

  EnvState st;
  float slot[8];
  if (live) {
    load_state(S, e, st);
    load_slot(S, e, sub, slot);
  }
  for (int t = 0; t < A.n_steps; ++t) {
    const bool run = live && !(st.flags & kFlagFinished);
    if (__all_sync(0xFFFFFFFFu, run)) step_once<MODE>(S, A, st, slot, e, i, sub, 0xFFFFFFFFu, t);   // the common case
    else if (live) step_once<MODE>(S, A, st, slot, e, i, sub, gmask, t);
  }
  if (live) {
    if (A.n_steps != 1) store_slot(S, e, sub, slot);
    store_state(S, e, st, sub);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreadsPerBlock, 4)
reset_kernel(const __grid_constant__ SimDev S, const int32_t *__restrict__ env_ids, int n, float *__restrict__ obs,
             int64_t obs_stride) {
  const int i = blockIdx.x * kEnvsPerBlock + (threadIdx.x >> 3);
  if (i >= n) return;
  const int sub = threadIdx.x & 7;
  const unsigned gmask = group_mask();
  const int e = env_ids ? __ldg(env_ids + i) : i;
  if (e < 0 || e >= S.n_envs) {
    if (sub == 0) atomicExch(S.error_flag, 2);
    return;
  }
  EnvState st;
  load_state(S, e, st);
  reset_episode(S, st);
  float slot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (MODE != MANSY_OBS_NONE && obs) emit_obs<MODE>(S, st, slot, sub, gmask, obs + (size_t)i * obs_stride);
  store_state(S, e, st, sub);
}

// ------------------------------------------------------------------------------------------
// MPC expert: ExpertEnv.choose_action, bitrate_selection/envs/expert_env.py:358-422
// ------------------------------------------------------------------------------------------
// One CTA per environment.  The reference scores all 15^horizon action sequences of the next H = min(horizon, chunks
// left) chunks with virtual downloads from the environment's trace / buffer state (simulator.py:125-144) and the
// QoE of utils/qoe.py:50-60 on cached per-(chunk, action) statistics (expert_env.py:121-160), and returns the first
// action of the first sequence with the largest sum.  Here: phase 0 computes the <= 6 x 15 statistics the search
// needs (tiles allocated from the predicted viewport, quality over the actual viewport -- the same numbers step_env
// produces for that chunk and action); phase 1 gives every thread prefixes (a_0 .. a_{H-2}) of the sequence tree,
// each followed by its 15 leaves, so a leaf costs 1 + (H-1)/15 virtual downloads instead of H; a CTA-wide arg-max
// with the reference's tie rule (lowest sequence index, digit t = action of step t) picks the action.
// float64 chain like step_env (the pinned numpy 1.24 stack).
constexpr int kExpertThreads = 256;
constexpr int kExpertMaxHorizon = 6;

struct ExpertProf { double size, vq, dev; };

__device__ __forceinline__ void expert_virtual_step(const SimDev &S, const ExpertProf &p, double sm, const double *__restrict__ tr,
                                                    int tlen, int &cur_idx, double &cur_time, double &buf, double &prev_vq,
                                                    bool first, double w0, double w1, double w2, double &sum, bool &ok) {
  const double dl = trace_download(p.size, [tr](int i) { return __ldg(tr + i); }, tlen, cur_idx, cur_time, ok);
  const double rebuf = buffer_push(buf, S.chunk_length, dl);
  const QoE r = qoe_from_sums(p.vq, p.dev, sm, rebuf, first, prev_vq, w0, w1, w2, S.max_quality);
  sum = dadd(sum, r.qoe);
}

__global__ void __launch_bounds__(kExpertThreads, 2) expert_mpc_kernel(const SimDev S, int horizon, int32_t *__restrict__ actions,
                                                                    double *__restrict__ best_value) {
  __shared__ ExpertProf prof[kExpertMaxHorizon][16];
  __shared__ double s_sm[kExpertMaxHorizon];
  __shared__ double red_val[kExpertThreads / 32];
  __shared__ int red_idx[kExpertThreads / 32];
  const int e = blockIdx.x;
  const EnvState st = S.state[e];
  const int H = min(horizon, st.end_chunk - st.next_chunk + 1);
  if ((st.flags & kFlagFinished) || H <= 0) {     // no chunk left: every sequence scores 0, the reference returns action 0
    if (threadIdx.x == 0) {
      actions[e] = 0;
      if (best_value) best_value[e] = 0.0;
    }
    return;
  }
  // ---- phase 0: chunk_pred_sizes / viewport qualities / intra variance of (chunk next + t, action a) ----
  for (int k = threadIdx.x; k < H * kActions; k += kExpertThreads) {
    const int t = k / kActions, a = k % kActions;
    const int c = st.next_chunk + t;
    const size_t vi = (size_t)st.pair * S.n_vp_chunks + (c - st.start_chunk);
    const uint64_t gt = __ldg(S.vp_gt + vi);
    const size_t tab = ((size_t)st.video * S.n_chunks + c) * kTableRow;
    int rin, rout;
    action_to_rates(a, rin, rout);
    const uint32_t vtab = (S.lut[rout] & ~7u) | (uint32_t)rin;
    int sz = 0;
    double mq = 0.0;
    for (int row = 0; row < 8; ++row) {
      const uint32_t scales = __ldg(S.vp_scale + vi * 8 + row);
      for (int i = 0; i < 8; ++i) {
        const int tile = row * 8 + i;
        const int ver = (int)((vtab >> (3u * ((scales >> (4 * i)) & 7u))) & 7u);
        sz += __ldg(S.size + tab + ver * kTiles + tile);
        if ((gt >> tile) & 1ULL) mq = dadd(mq, (double)__ldg(S.quality + tab + ver * kTiles + tile));
      }
    }
    const double sm = (double)__popcll(gt);
    const double vq = ddiv(mq, sm);
    const float vq32 = (float)vq;
    double dev = 0.0;
    for (int row = 0; row < 8; ++row) {
      const uint32_t scales = __ldg(S.vp_scale + vi * 8 + row);
      for (int i = 0; i < 8; ++i) {
        const int tile = row * 8 + i;
        if (!((gt >> tile) & 1ULL)) continue;
        const int ver = (int)((vtab >> (3u * ((scales >> (4 * i)) & 7u))) & 7u);
        dev = dadd(dev, (double)fabsf(fsub(__ldg(S.quality + tab + ver * kTiles + tile), vq32)));
      }
    }
    prof[t][a].size = (double)sz;
    prof[t][a].vq = vq;
    prof[t][a].dev = dev;
    if (a == 0) s_sm[t] = sm;
  }
  __syncthreads();
  // ---- phase 1: sequence tree ----
  const double *tr = S.trace + (size_t)st.trace * S.trace_stride;
  const int tlen = __ldg(S.trace_len + st.trace);
  const double w0 = (double)st.w0, w1 = (double)st.w1, w2 = (double)st.w2;
  int n_prefix = 1;
  for (int t = 0; t < H - 1; ++t) n_prefix *= kActions;
  double bv = -INFINITY;
  int bi = 0x7FFFFFFF;
  bool ok = true;
  for (int p = threadIdx.x; p < n_prefix; p += kExpertThreads) {
    int cur_idx = st.cur_idx;
    double cur_time = st.cur_time, buf = st.buf, prev = st.prev_vq, sum = 0.0;
    bool first = st.ep_step == 0;
    int tmp = p;
    for (int t = 0; t < H - 1; ++t) {
      const int a = tmp % kActions;
      tmp /= kActions;
      expert_virtual_step(S, prof[t][a], s_sm[t], tr, tlen, cur_idx, cur_time, buf, prev, first, w0, w1, w2, sum, ok);
      first = false;
    }
    for (int a = 0; a < kActions; ++a) {
      int li = cur_idx;
      double lt = cur_time, lb = buf, lp = prev, ls = sum;
      expert_virtual_step(S, prof[H - 1][a], s_sm[H - 1], tr, tlen, li, lt, lb, lp, first, w0, w1, w2, ls, ok);
      const int idx = p + a * n_prefix;
      if (ls > bv || (ls == bv && idx < bi)) { bv = ls; bi = idx; }
    }
  }
  if (!ok) atomicExch(S.error_flag, 1);
  // ---- arg-max, lowest index on ties (expert_env.py:408-410: strict '<' keeps the first maximum) ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { red_val[threadIdx.x >> 5] = bv; red_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kExpertThreads / 32; ++w)
      if (red_val[w] > bv || (red_val[w] == bv && red_idx[w] < bi)) { bv = red_val[w]; bi = red_idx[w]; }
    actions[e] = bi % kActions;                   // rates2action(action2rates(a)) == a (utils/common.py:101-139)
    if (best_value) best_value[e] = bv;
  }
}

// Fills the (viewport pair, chunk, action) outcome table with the gather form of step_env (chunk_parts_gather), one
// 8-lane group per entry: what a step then reads back is bit-identical to what it would have computed.
__global__ void __launch_bounds__(kThreadsPerBlock) outcome_build_kernel(const SimDev S, ChunkOutcome *__restrict__ out, long long total) {
  const long long g = (long long)blockIdx.x * kEnvsPerBlock + (threadIdx.x >> 3);
  if (g >= total) return;
  const int sub = threadIdx.x & 7;
  const unsigned gmask = group_mask();
  const int a = (int)(g % kOutcomeActions);
  const long long vi = g / kOutcomeActions;
  const int j = (int)(vi % S.n_vp_chunks), pair = (int)(vi / S.n_vp_chunks);
  const int video = pair / S.n_users;
  const int c = __ldg(S.vp_start + pair) + j;
  const int end = min(__ldg(S.vp_end + pair), __ldg(S.video_time + video) - 1);
  ChunkOutcome o;
  o.q1 = 0.0; o.intra = 0.0; o.size = 0; o.pad0 = 0; o.pad1 = 0;
  if (c >= 0 && c <= end && c < S.n_chunks) {           // chunks an episode of this pair can download (simulator.py:41-45)
    int rin, rout;
    action_to_rates(a, rin, rout);                       // a == 15: out of the table -> (0, 0)
    const ChunkParts cp = chunk_parts_gather(S, video, c, __ldg(S.vp_gt + vi), __ldg(S.vp_scale + vi * 8 + sub), rin, rout, sub,
                                             gmask, nullptr);
    o.q1 = cp.q1; o.intra = cp.intra; o.size = cp.sz;
  }
  if (sub == 0) out[g] = o;
}

__global__ void copy_i32_kernel(int32_t *__restrict__ dst, const int32_t *__restrict__ src, int32_t n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
  __threadfence_system();
}

// reward / done / log-prob / value of one step, stored straight into the caller's mapped pinned buffers: four 4-16 KB
// cudaMemcpyAsync calls cost the device-to-host copy engine more idle time between the 12.9 MB observation slabs than
// the bytes they move
__global__ void store_step_scalars_kernel(float *__restrict__ reward_h, uint8_t *__restrict__ done_h, float *__restrict__ logp_h,
                                          float *__restrict__ value_h, const float *__restrict__ reward, const uint8_t *__restrict__ done,
                                          const float *__restrict__ logp, const float *__restrict__ value, int32_t n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    reward_h[i] = reward[i];
    done_h[i] = done[i];
    logp_h[i] = logp[i];
    value_h[i] = value[i];
  }
  __threadfence_system();
}

__global__ void seed_kernel(const SimDev S, int32_t seed, int full) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S.n_envs) return;
  // mansy_env.py:253-256 with the vector env handing env k the seed `seed + k`: seed() only moves worker_id, a
  // running episode is left alone (`full` = the state wipe of mansy_create)
  long long wid = ((long long)seed + S.env_offset + e) % S.worker_num;
  if (wid < 0) wid += S.worker_num;
  if (!full) {
    S.state[e].cursor = (int32_t)wid;
    return;
  }
  EnvState st;
  memset(&st, 0, sizeof(st));
  st.cursor = (int32_t)wid;
  st.flags = kFlagFinished | (kNoAction << 8);     // must be reset before the first step
  st.end_chunk = S.startup_download + 1;
  st.next_chunk = S.startup_download + 1;
  st.start_chunk = 0;
  S.state[e] = st;
}

__global__ void stats_clear_kernel(double *stats, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stats[i] = 0.0;
}

// viewport_prediction/predict.py:33-48: one thread per chunk.
__global__ void viewport_tiles_kernel(const float *__restrict__ gt_xy, const float *__restrict__ pred_xy, int64_t n,
                                      int points, int width, int height, int fov_w, int fov_h,
                                      uint64_t *__restrict__ gt_mask, uint64_t *__restrict__ pred_mask,
                                      double *__restrict__ acc, int32_t *__restrict__ invalid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool valid = true;
  uint64_t g = 0, p = 0;
  const float2 *gp = reinterpret_cast<const float2 *>(gt_xy) + i * points;
  for (int j = 0; j < points; ++j) {
    const float2 v = __ldg(gp + j);
    g |= fov_tile_mask(centre_to_pixel(v.x, width), centre_to_pixel(v.y, height), width, height, fov_w, fov_h, valid);
  }
  gt_mask[i] = g;
  if (pred_xy) {
    const float2 *pp = reinterpret_cast<const float2 *>(pred_xy) + i * points;
    for (int j = 0; j < points; ++j) {
      const float2 v = __ldg(pp + j);
      p |= fov_tile_mask(centre_to_pixel(v.x, width), centre_to_pixel(v.y, height), width, height, fov_w, fov_h, valid);
    }
    pred_mask[i] = p;
    acc[i] = ddiv((double)__popcll(g & p), (double)__popcll(g | p));   // IoU, predict.py:46
  }
  if (!valid && invalid) atomicExch(invalid, 1);
}

// utils/common.py:101-119,142-193 for standalone masks: one thread per (mask, action).
__global__ void allocate_versions_kernel(const uint64_t *__restrict__ masks, const int32_t *__restrict__ actions,
                                         int64_t n, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t l4,
                                         uint8_t *__restrict__ versions) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t lut[5] = {l0, l1, l2, l3, l4};
  int rin, rout;
  action_to_rates(__ldg(actions + i), rin, rout);
  const TileScaleMasks dm = tile_scale_masks(__ldg(masks + i));
  const uint32_t lutw = lut[rout];
  uint4 *dst = reinterpret_cast<uint4 *>(versions + i * kTiles);
#pragma unroll
  for (int qd = 0; qd < 4; ++qd) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      const int t = qd * 16 + b;
      w[b >> 2] |= (uint32_t)tile_version(lutw, rin, tile_scale(dm, t)) << (8 * (b & 3));
    }
    dst[qd] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace mansy

// ------------------------------------------------------------------------------------------
// host side: handle + C ABI
// ------------------------------------------------------------------------------------------
using namespace mansy;

namespace mansy {
// tensor-core policy launch (mansy_policy_tc.cu); mansy_policy_t is opaque here
int policy_device_of(mansy_policy_t p);     // mansy_policy.cu
int policy_forward_tc_launch(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                             float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                             int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, int64_t *timeline_dev,
                             int32_t timeline_cta, bool pdl, void *stream, const SimDev *memo_sim);
int rollout_fused_launch(mansy_policy_t p, const SimDev &S, const mansy_rollout_t *b, int32_t n_steps, int64_t t0,
                         uint64_t seed, void *stream, int *launched);
}  // namespace mansy

struct mansy_sim {
  SimDev dev;
  int device = 0;
  ChunkOutcome *outcome = nullptr;    // the table dev.outcome points at when enabled (mansy_set_outcome_table)
  std::vector<void *> allocs;
  // lazily allocated device staging for the host-buffer entry points
  int32_t *stage_actions = nullptr;
  float *stage_obs = nullptr;
  float *stage_reward = nullptr;
  uint8_t *stage_done = nullptr;
  int64_t obs_stride = 0;
  // events of the last instrumented mansy_rollout_policy call: [3 * steps] = before policy, after policy, after step
  std::vector<cudaEvent_t> events;
  int timed_steps = 0;
  // host-storage rollout: result copies run on their own stream, overlapped with the next step's kernels
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> step_done, copy_done;     // rings of kCopyRing events
};
constexpr int kCopyRing = 64;

namespace mansy {
const SimDev *sim_dev_of(mansy_handle_t h) { return &h->dev; }    // mansy_peer.cu packs the statistics rows
int sim_device_of(mansy_handle_t h) { return h->device; }
}

namespace {

template <typename T>
int upload(mansy_sim *h, const T *host, size_t count, const T **dev_out) {
  void *d = nullptr;
  if (cudaMalloc(&d, count * sizeof(T) + 16) != cudaSuccess) return set_error(MANSY_E_NOMEM, "cudaMalloc failed (table)");
  h->allocs.push_back(d);
  MANSY_CUDA(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
  *dev_out = static_cast<const T *>(d);
  return MANSY_OK;
}

template <typename T>
int dev_alloc(mansy_sim *h, size_t count, T **out) {
  void *d = nullptr;
  if (cudaMalloc(&d, count * sizeof(T) + 16) != cudaSuccess) return set_error(MANSY_E_NOMEM, "cudaMalloc failed (state)");
  h->allocs.push_back(d);
  *out = static_cast<T *>(d);
  return MANSY_OK;
}

int check_tables(const mansy_tables_t *t) {
  if (!t) return set_error(MANSY_E_INVALID, "tables is NULL");
  if (!t->size || !t->quality || !t->video_time || !t->vp_gt || !t->vp_pred || !t->vp_acc || !t->vp_start ||
      !t->vp_end || !t->trace || !t->trace_len || !t->qoe_w || !t->samples)
    return set_error(MANSY_E_INVALID, "a table pointer is NULL");
  if (t->n_videos < 1 || t->n_chunks < 1 || t->n_users < 1 || t->n_vp_chunks < 1 || t->n_traces < 1 ||
      t->trace_stride < 1 || t->n_qoe < 1 || t->n_samples < 1)
    return set_error(MANSY_E_INVALID, "a table dimension is < 1");
  return MANSY_OK;
}

inline int grid_for(int n) { return (n + kEnvsPerBlock - 1) / kEnvsPerBlock; }

int obs_stride_for(int mode) {
  return mode == MANSY_OBS_MANSY ? MANSY_OBS_MANSY_STRIDE : (mode == MANSY_OBS_SIMPLE ? MANSY_OBS_SIMPLE_STRIDE : 0);
}

int check_out(const mansy_sim *h, const mansy_out_t *out) {
  if (!out) return set_error(MANSY_E_INVALID, "out is NULL");
  if (out->obs) {
    if (h->dev.obs_mode == MANSY_OBS_NONE) return set_error(MANSY_E_INVALID, "handle was created with MANSY_OBS_NONE");
    if (out->obs_stride < obs_stride_for(h->dev.obs_mode) || (out->obs_stride & 3))
      return set_error(MANSY_E_INVALID, "obs_stride too small or not a multiple of 4 floats");
    if (reinterpret_cast<uintptr_t>(out->obs) & 15) return set_error(MANSY_E_INVALID, "obs must be 16-byte aligned");
  }
  if (out->aux && (reinterpret_cast<uintptr_t>(out->aux) & 15)) return set_error(MANSY_E_INVALID, "aux must be 16-byte aligned");
  if (out->tile_versions && (reinterpret_cast<uintptr_t>(out->tile_versions) & 7))
    return set_error(MANSY_E_INVALID, "tile_versions must be 8-byte aligned");
  return MANSY_OK;
}

int launch_step(mansy_sim *h, const StepArgs &a, cudaStream_t s, bool pdl = false) {
  const int grid = grid_for(a.n);
  if (grid == 0) return MANSY_OK;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kThreadsPerBlock, 1, 1);
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  switch (h->dev.obs_mode) {
    case MANSY_OBS_MANSY: MANSY_CUDA(cudaLaunchKernelEx(&cfg, step_kernel<MANSY_OBS_MANSY>, h->dev, a)); break;
    case MANSY_OBS_SIMPLE: MANSY_CUDA(cudaLaunchKernelEx(&cfg, step_kernel<MANSY_OBS_SIMPLE>, h->dev, a)); break;
    default: MANSY_CUDA(cudaLaunchKernelEx(&cfg, step_kernel<MANSY_OBS_NONE>, h->dev, a)); break;
  }
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

int ensure_staging(mansy_sim *h) {
  if (h->stage_actions) return MANSY_OK;
  const size_t n = (size_t)h->dev.n_envs;
  h->obs_stride = obs_stride_for(h->dev.obs_mode);
  int rc;
  if ((rc = dev_alloc(h, n, &h->stage_actions))) return rc;
  if ((rc = dev_alloc(h, n, &h->stage_reward))) return rc;
  if ((rc = dev_alloc(h, n, &h->stage_done))) return rc;
  if (h->obs_stride && (rc = dev_alloc(h, n * (size_t)h->obs_stride, &h->stage_obs))) return rc;
  return MANSY_OK;
}

}  // namespace

static int launch_seed(mansy_handle_t h, int32_t seed, int full, void *stream);

extern "C" {

const char *mansy_last_error(void) { return g_last_error.c_str(); }
int mansy_abi_version(void) { return MANSY_ABI_VERSION; }
int64_t mansy_kernel_launches(void) { return g_launches.load(); }

int mansy_create(const mansy_tables_t *t, const mansy_cfg_t *cfg, int device, mansy_handle_t *out) {
  if (!out) return set_error(MANSY_E_INVALID, "out is NULL");
  *out = nullptr;
  int rc = check_tables(t);
  if (rc) return rc;
  if (!cfg) return set_error(MANSY_E_INVALID, "cfg is NULL");
  if (cfg->n_envs < 1 || cfg->worker_num < 1 || cfg->env_offset < 0)
    return set_error(MANSY_E_INVALID, "n_envs / worker_num must be >= 1 and env_offset >= 0");
  if (cfg->obs_mode < MANSY_OBS_NONE || cfg->obs_mode > MANSY_OBS_SIMPLE) return set_error(MANSY_E_INVALID, "bad obs_mode");
  if (cfg->reward_mode != MANSY_REWARD_QOE && cfg->reward_mode != MANSY_REWARD_QOE_NORM)
    return set_error(MANSY_E_INVALID, "bad reward_mode");
  for (int i = 0; i < 5; ++i)
    if (cfg->video_rates[i] <= 0 || (i && cfg->video_rates[i] < cfg->video_rates[i - 1]))
      return set_error(MANSY_E_INVALID, "video_rates must be positive and ascending");
  if (cfg->startup_download < 0 || cfg->chunk_length < 1 || cfg->max_size < 1 || cfg->max_throughput < 1)
    return set_error(MANSY_E_INVALID, "bad streaming constants");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return set_error(MANSY_E_CUDA, "no CUDA device available: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(MANSY_E_INVALID, "bad device index");
  DeviceScope dscope(device);
  MANSY_CUDA(dscope.err);

  mansy_sim *h = new (std::nothrow) mansy_sim();
  if (!h) return set_error(MANSY_E_NOMEM, "out of host memory");
  h->device = device;
  SimDev &d = h->dev;
  memset(&d, 0, sizeof(d));
  static std::atomic<uint64_t> next_uid{1};
  d.uid = next_uid.fetch_add(1);
  const size_t n_tab = (size_t)t->n_videos * t->n_chunks * kTableRow;
  const size_t n_vp = (size_t)t->n_videos * t->n_users * t->n_vp_chunks;
  const size_t n_pairs = (size_t)t->n_videos * t->n_users;

  // host-side validation that the kernels rely on (the reference would raise / loop forever)
  for (size_t p = 0; p < n_pairs && !rc; ++p) {
    const int v = (int)(p / t->n_users);
    const int end = t->vp_end[p] < t->video_time[v] - 1 ? t->vp_end[p] : t->video_time[v] - 1;
    if (t->vp_start[p] > cfg->startup_download + 1) rc = set_error(MANSY_E_INVALID, "viewport trace starts after the first simulated chunk (simulator.py:44)");
    else if (end < cfg->startup_download + 1) rc = set_error(MANSY_E_INVALID, "an episode would have no chunk to simulate");
    else if (end >= t->n_chunks) rc = set_error(MANSY_E_INVALID, "end chunk beyond the size table");
    else if (t->vp_end[p] - t->vp_start[p] + 1 > t->n_vp_chunks) rc = set_error(MANSY_E_INVALID, "viewport chunk range exceeds n_vp_chunks");
  }
  for (int k = 0; k < t->n_traces && !rc; ++k) {
    if (t->trace_len[k] < 1 || t->trace_len[k] > t->trace_stride) { rc = set_error(MANSY_E_INVALID, "bad trace_len"); break; }
    bool pos = false;
    for (int i = 0; i < t->trace_len[k]; ++i) {
      const double x = t->trace[(size_t)k * t->trace_stride + i];
      if (!(x >= 0.0) || x > 1e300) { rc = set_error(MANSY_E_INVALID, "negative / non-finite throughput"); break; }
      pos |= x > 0.0;
    }
    if (!rc && !pos) rc = set_error(MANSY_E_INVALID, "a trace has no positive-throughput second (network.py:24 would never return)");
  }
  for (int s = 0; s < t->n_samples && !rc; ++s) {
    const int32_t *q = t->samples + 4 * (size_t)s;
    if (q[0] < 0 || q[0] >= t->n_videos || q[1] < 0 || q[1] >= t->n_users || q[2] < 0 || q[2] >= t->n_traces ||
        q[3] < 0 || q[3] >= t->n_qoe)
      rc = set_error(MANSY_E_INVALID, "sample index out of range");
  }
  if (rc) { delete h; return rc; }

  // derived tables: normalised observation rows (utils/common.py:40-47), IEEE float32 division
  std::vector<float> size_norm(n_tab), qual_norm(n_tab);
  const float fmax_size = (float)cfg->max_size, fmax_q = (float)cfg->video_rates[4];
  for (size_t i = 0; i < n_tab; ++i) {
    if (t->size[i] <= 0) { delete h; return set_error(MANSY_E_INVALID, "tile sizes must be positive"); }
    size_norm[i] = fdiv((float)t->size[i], fmax_size);
    qual_norm[i] = fdiv(t->quality[i], fmax_q);
  }

#define MANSY_TRY(x) do { rc = (x); if (rc) { mansy_destroy(h); return rc; } } while (0)
  MANSY_TRY(upload(h, t->size, n_tab, &d.size));
  MANSY_TRY(upload(h, t->quality, n_tab, &d.quality));
  MANSY_TRY(upload(h, size_norm.data(), n_tab, &d.size_norm));
  MANSY_TRY(upload(h, qual_norm.data(), n_tab, &d.qual_norm));
  {
    std::vector<float> merged(2 * n_tab);
    for (size_t r = 0; r < n_tab / kTableRow; ++r) {
      memcpy(&merged[r * 2 * kTableRow], &size_norm[r * kTableRow], kTableRow * sizeof(float));
      memcpy(&merged[r * 2 * kTableRow + kTableRow], &qual_norm[r * kTableRow], kTableRow * sizeof(float));
    }
    MANSY_TRY(upload(h, merged.data(), merged.size(), &d.obs_tab));
  }
  MANSY_TRY(upload(h, t->video_time, (size_t)t->n_videos, &d.video_time));
  MANSY_TRY(upload(h, t->vp_gt, n_vp, &d.vp_gt));
  MANSY_TRY(upload(h, t->vp_pred, n_vp, &d.vp_pred));
  {
    // allocate_tile_rates' scale map (utils/common.py:142-168) once per (viewport pair, chunk) instead of once per
    // environment step: 64 toroidal distances, 4 bits each, one 32-bit word per tile row
    std::vector<uint32_t> scale(n_vp * 8);
    for (size_t v = 0; v < n_vp; ++v) {
      const TileScaleMasks m = tile_scale_masks(t->vp_pred[v]);
      for (int row = 0; row < 8; ++row) {
        uint32_t w = 0;
        for (int i = 0; i < 8; ++i) w |= (uint32_t)tile_scale(m, row * 8 + i) << (4 * i);
        scale[v * 8 + row] = w;
      }
    }
    MANSY_TRY(upload(h, scale.data(), scale.size(), &d.vp_scale));
  }
  MANSY_TRY(upload(h, t->vp_acc, n_vp, &d.vp_acc));
  MANSY_TRY(upload(h, t->vp_start, n_pairs, &d.vp_start));
  MANSY_TRY(upload(h, t->vp_end, n_pairs, &d.vp_end));
  {
    // device copy of the traces with 8 wrap-around entries appended to every row (entry[len + m] = entry[m % len]),
    // so that the 8 lanes of an environment can fetch the window cur_idx .. cur_idx + 7 with one load each and no
    // modulo (network.py:29 wraps the index)
    const int stride_dev = t->trace_stride + kTraceWindow;
    std::vector<double> padded((size_t)t->n_traces * stride_dev, 0.0);
    for (int k = 0; k < t->n_traces; ++k) {
      const double *src = t->trace + (size_t)k * t->trace_stride;
      double *dst = padded.data() + (size_t)k * stride_dev;
      const int len = t->trace_len[k];
      for (int i = 0; i < len; ++i) dst[i] = src[i];
      for (int m = 0; m < kTraceWindow; ++m) dst[len + m] = src[m % len];
    }
    MANSY_TRY(upload(h, padded.data(), padded.size(), &d.trace));
    d.trace_stride = stride_dev;
    for (double &x : padded) x = 1.0 / x;               // correctly rounded reciprocals for ddiv_rcp (1 / 0 = inf is never used)
    MANSY_TRY(upload(h, padded.data(), padded.size(), &d.trace_rcp));
  }
  MANSY_TRY(upload(h, t->trace_len, (size_t)t->n_traces, &d.trace_len));
  MANSY_TRY(upload(h, t->qoe_w, (size_t)t->n_qoe * 3, &d.qoe_w));
  MANSY_TRY(upload(h, t->samples, (size_t)t->n_samples * 4, &d.samples));
  {
    std::vector<EpisodeInit> init((size_t)t->n_samples);
    for (int s = 0; s < t->n_samples; ++s) {
      const int32_t *q = t->samples + 4 * (size_t)s;
      EpisodeInit &e = init[(size_t)s];
      memset(&e, 0, sizeof(e));
      e.video = q[0];
      e.pair = q[0] * t->n_users + q[1];
      e.trace = q[2];
      e.w0 = t->qoe_w[q[3] * 3 + 0]; e.w1 = t->qoe_w[q[3] * 3 + 1]; e.w2 = t->qoe_w[q[3] * 3 + 2];
      e.start_chunk = t->vp_start[e.pair];
      const int vend = t->vp_end[e.pair], tend = t->video_time[e.video] - 1;
      e.end_chunk = vend < tend ? vend : tend;
      const int first = cfg->startup_download + 1 < e.end_chunk ? cfg->startup_download + 1 : e.end_chunk;
      e.first_pred = t->vp_pred[(size_t)e.pair * t->n_vp_chunks + (first - e.start_chunk)];
    }
    MANSY_TRY(upload(h, init.data(), init.size(), &d.ep_init));
  }
  d.n_videos = t->n_videos; d.n_chunks = t->n_chunks; d.n_users = t->n_users; d.n_vp_chunks = t->n_vp_chunks;
  d.n_traces = t->n_traces; d.n_qoe = t->n_qoe; d.n_samples = t->n_samples;

  const size_t n = (size_t)cfg->n_envs;
  MANSY_TRY(dev_alloc(h, n, &d.state));
  MANSY_TRY(dev_alloc(h, n * kHistFloatsPerEnv, &d.hist));
  MANSY_TRY(dev_alloc(h, n * MANSY_STATS_DOUBLES, &d.stats));
  MANSY_TRY(dev_alloc(h, (size_t)4, &d.error_flag));
  d.n_envs = cfg->n_envs; d.env_offset = cfg->env_offset; d.worker_num = cfg->worker_num;
  d.obs_mode = cfg->obs_mode; d.reward_mode = cfg->reward_mode;
  d.startup_download = cfg->startup_download;
  d.chunk_length = (double)cfg->chunk_length;
  d.max_quality = (double)cfg->video_rates[4];
  d.max_throughput = (double)cfg->max_throughput;
  d.startup_d = (double)cfg->startup_download;
  d.startup_f = (float)cfg->startup_download;
  d.rcp_max_quality = 1.0 / d.max_quality;
  d.rcp_max_throughput = 1.0 / d.max_throughput;
  d.rcp_startup_d = 1.0 / d.startup_d;
  for (int k = 0; k <= kTiles; ++k) d.rcp_count[k] = 1.0 / (double)k;
  for (int i = 0; i < kRates; ++i) {
    d.rate_norm_hist[i] = (float)((double)cfg->video_rates[i] / (double)cfg->video_rates[4]);
    d.rate_norm_f32[i] = fdiv((float)cfg->video_rates[i], (float)cfg->video_rates[4]);
  }
  build_rate_lut(cfg->video_rates, d.lut);
  {
    cudaError_t e1 = cudaMemset(d.hist, 0, n * kHistFloatsPerEnv * sizeof(float));
    cudaError_t e2 = cudaMemset(d.stats, 0, n * MANSY_STATS_DOUBLES * sizeof(double));
    cudaError_t e3 = cudaMemset(d.error_flag, 0, 16);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { mansy_destroy(h); return set_error(MANSY_E_CUDA, "cudaMemset failed"); }
  }
  {
    // (viewport pair, chunk, action) outcome table: 32 B x 16 actions per (pair, chunk) -- 40 MB at 1 440 pairs x 54
    // chunks; skipped above 1 GiB (steps then gather) or with MANSY_NO_OUTCOME_TABLE=1
    const long long total = (long long)n_vp * kOutcomeActions;
    const char *off = getenv("MANSY_NO_OUTCOME_TABLE");
    if (!(off && off[0] == '1') && total * (long long)sizeof(ChunkOutcome) <= (1LL << 30)) {
      MANSY_TRY(dev_alloc(h, (size_t)total, &h->outcome));
      const long long grid = (total + kEnvsPerBlock - 1) / kEnvsPerBlock;
      outcome_build_kernel<<<(unsigned)grid, kThreadsPerBlock>>>(d, h->outcome, total);
      count_launch();
      if (cudaGetLastError() != cudaSuccess) { mansy_destroy(h); return set_error(MANSY_E_CUDA, "outcome_build_kernel launch failed"); }
      d.outcome = h->outcome;
    }
  }
#undef MANSY_TRY
  rc = launch_seed(h, cfg->seed, 1, nullptr);
  if (rc) { mansy_destroy(h); return rc; }
  cudaError_t e = cudaStreamSynchronize(nullptr);
  if (e != cudaSuccess) { mansy_destroy(h); return set_error(MANSY_E_CUDA, cudaGetErrorString(e)); }
  *out = h;
  return MANSY_OK;
}

int mansy_destroy(mansy_handle_t h) {
  if (!h) return MANSY_OK;
  DeviceScope dscope(h->device);
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  for (cudaEvent_t e : h->step_done) cudaEventDestroy(e);
  for (cudaEvent_t e : h->copy_done) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (void *p : h->allocs) cudaFree(p);
  delete h;
  return MANSY_OK;
}

}  // extern "C"

static int launch_seed(mansy_handle_t h, int32_t seed, int full, void *stream) {
  const int threads = 256, grid = (h->dev.n_envs + threads - 1) / threads;
  seed_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->dev, seed, full);
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

extern "C" {

int mansy_seed(mansy_handle_t h, int32_t seed, void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  return launch_seed(h, seed, 0, stream);
}

int mansy_reset(mansy_handle_t h, const int32_t *env_ids_dev, int32_t n, float *obs_dev, int64_t obs_stride,
                void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  if (n < 0 || (!env_ids_dev && n != h->dev.n_envs))
    return set_error(MANSY_E_INVALID, "n must equal n_envs when env_ids is NULL");
  if (n == 0) return MANSY_OK;
  if (obs_dev) {
    mansy_out_t o; memset(&o, 0, sizeof(o)); o.obs = obs_dev; o.obs_stride = obs_stride;
    int rc = check_out(h, &o);
    if (rc) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n);
  switch (h->dev.obs_mode) {
    case MANSY_OBS_MANSY: reset_kernel<MANSY_OBS_MANSY><<<grid, kThreadsPerBlock, 0, s>>>(h->dev, env_ids_dev, n, obs_dev, obs_stride); break;
    case MANSY_OBS_SIMPLE: reset_kernel<MANSY_OBS_SIMPLE><<<grid, kThreadsPerBlock, 0, s>>>(h->dev, env_ids_dev, n, obs_dev, obs_stride); break;
    default: reset_kernel<MANSY_OBS_NONE><<<grid, kThreadsPerBlock, 0, s>>>(h->dev, env_ids_dev, n, obs_dev, obs_stride); break;
  }
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

int mansy_step(mansy_handle_t h, const int32_t *actions_dev, const int32_t *env_ids_dev, int32_t n, int32_t auto_reset,
               const mansy_out_t *out, void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  if (!actions_dev) return set_error(MANSY_E_INVALID, "actions is NULL");
  if (n < 0 || (!env_ids_dev && n != h->dev.n_envs))
    return set_error(MANSY_E_INVALID, "n must equal n_envs when env_ids is NULL");
  int rc = check_out(h, out);
  if (rc) return rc;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.actions = actions_dev; a.env_ids = env_ids_dev; a.n = n; a.auto_reset = auto_reset; a.action_mode = 0;
  a.n_steps = 1; a.out = *out;
  return launch_step(h, a, static_cast<cudaStream_t>(stream));
}

int mansy_rollout_random(mansy_handle_t h, int32_t n_steps, uint64_t seed, int64_t step0, int32_t per_step_outputs,
                         const mansy_out_t *out, void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  if (n_steps < 1) return set_error(MANSY_E_INVALID, "n_steps must be >= 1");
  int rc = check_out(h, out);
  if (rc) return rc;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n = h->dev.n_envs; a.auto_reset = 1; a.action_mode = 1; a.n_steps = n_steps; a.seed = seed; a.step0 = step0;
  a.rows_per_step = per_step_outputs ? h->dev.n_envs : 0;
  a.out = *out;
  return launch_step(h, a, static_cast<cudaStream_t>(stream));
}

int mansy_step_host(mansy_handle_t h, const int32_t *actions_host, int32_t auto_reset, float *obs_host,
                    float *reward_host, uint8_t *done_host, void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  if (!actions_host) return set_error(MANSY_E_INVALID, "actions is NULL");
  int rc = ensure_staging(h);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)h->dev.n_envs;
  MANSY_CUDA(cudaMemcpyAsync(h->stage_actions, actions_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.actions = h->stage_actions; a.n = h->dev.n_envs; a.auto_reset = auto_reset; a.n_steps = 1;
  a.out.obs = obs_host ? h->stage_obs : nullptr; a.out.obs_stride = h->obs_stride;
  a.out.reward = h->stage_reward; a.out.done = h->stage_done;
  rc = launch_step(h, a, s);
  if (rc) return rc;
  if (obs_host && h->stage_obs)
    MANSY_CUDA(cudaMemcpyAsync(obs_host, h->stage_obs, n * (size_t)h->obs_stride * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (reward_host) MANSY_CUDA(cudaMemcpyAsync(reward_host, h->stage_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (done_host) MANSY_CUDA(cudaMemcpyAsync(done_host, h->stage_done, n, cudaMemcpyDeviceToHost, s));
  MANSY_CUDA(cudaStreamSynchronize(s));
  return MANSY_OK;
}

int mansy_reset_host(mansy_handle_t h, float *obs_host, void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  int rc = ensure_staging(h);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = mansy_reset(h, nullptr, h->dev.n_envs, obs_host ? h->stage_obs : nullptr, h->obs_stride, stream);
  if (rc) return rc;
  if (obs_host && h->stage_obs)
    MANSY_CUDA(cudaMemcpyAsync(obs_host, h->stage_obs, (size_t)h->dev.n_envs * h->obs_stride * sizeof(float),
                               cudaMemcpyDeviceToHost, s));
  MANSY_CUDA(cudaStreamSynchronize(s));
  return MANSY_OK;
}

int mansy_rollout_reserve_timing(mansy_handle_t h, int32_t n_steps) {
  if (!h || n_steps < 0) return set_error(MANSY_E_INVALID, "bad argument");
  while (h->events.size() < (size_t)3 * n_steps) {
    cudaEvent_t e;
    MANSY_CUDA(cudaEventCreate(&e));
    h->events.push_back(e);
  }
  return MANSY_OK;
}

int mansy_rollout_policy(mansy_handle_t h, mansy_policy_t p, const mansy_rollout_t *b, int32_t n_steps, int64_t t0,
                         uint64_t seed, int32_t flags, void *stream) {
  if (!h || !p || !b) return set_error(MANSY_E_INVALID, "NULL argument");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  if (n_steps < 0 || t0 < 0) return set_error(MANSY_E_INVALID, "n_steps / t0 must be >= 0");
  if (b->slabs < 2) return set_error(MANSY_E_INVALID, "a rollout needs at least 2 observation slabs");
  if (!b->obs || !b->actions || !b->logp || !b->value || !b->reward || !b->done || !b->logits)
    return set_error(MANSY_E_INVALID, "a rollout buffer pointer is NULL");
  if (h->dev.obs_mode == MANSY_OBS_NONE) return set_error(MANSY_E_INVALID, "the policy consumes observation rows");
  if (policy_device_of(p) != h->device) return set_error(MANSY_E_INVALID, "policy and simulator live on different devices");
  const size_t n = (size_t)h->dev.n_envs;
  const bool timed = (flags & MANSY_ROLLOUT_TIME_KERNELS) != 0;
  // back-to-back launches overlap their scheduling / prologues (programmatic dependent launch) unless events
  // are recorded between them (a timed rollout measures the kernels one by one) or the caller opts out
  const bool pdl = !timed && !(flags & MANSY_ROLLOUT_NO_PDL) && !(flags & MANSY_ROLLOUT_FP32_POLICY);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (timed) {
    int rc = mansy_rollout_reserve_timing(h, n_steps);
    if (rc) return rc;
    h->timed_steps = n_steps;
  }
  const int is_probs = h->dev.obs_mode == MANSY_OBS_SIMPLE ? 1 : 0;
  if (!timed && !(flags & (MANSY_ROLLOUT_FP32_POLICY | MANSY_ROLLOUT_TWO_KERNELS)) && n_steps > 0) {
    // small batches: the whole loop is one launch of the fused cluster kernel (policy + sample + step per tile)
    int launched = 0;
    int rc = rollout_fused_launch(p, h->dev, b, n_steps, t0, seed, stream, &launched);
    if (rc) return rc;
    if (launched) return MANSY_OK;
  }
  for (int32_t k = 0; k < n_steps; ++k) {
    const int64_t t = t0 + k;
    const size_t cur = (size_t)(t % b->slabs), nxt = (size_t)((t + 1) % b->slabs);
    const float *obs = b->obs + cur * n * (size_t)b->obs_stride;
    int rc;
    if (timed) MANSY_CUDA(cudaEventRecord(h->events[3 * k], s));
    if (flags & MANSY_ROLLOUT_FP32_POLICY) {
      rc = mansy_policy_forward(p, obs, b->obs_stride, h->dev.n_envs, b->logits, b->value + cur * n, stream);
      if (!rc) rc = mansy_policy_sample(b->logits, h->dev.n_envs, is_probs, seed, t, h->dev.env_offset, b->actions + cur * n,
                                        b->logp + cur * n, stream);
    } else {
      rc = policy_forward_tc_launch(p, obs, b->obs_stride, h->dev.n_envs, b->logits, b->value + cur * n, b->actions + cur * n,
                                    b->logp + cur * n, seed, t, h->dev.env_offset, nullptr, nullptr, nullptr, 0, pdl, stream,
                                    getenv("MANSY_NO_POLICY_MEMO") ? nullptr : &h->dev);
    }
    if (rc) return rc;
    if (timed) MANSY_CUDA(cudaEventRecord(h->events[3 * k + 1], s));
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.actions = b->actions + cur * n; a.n = h->dev.n_envs; a.auto_reset = 1; a.n_steps = 1;
    a.out.obs = b->obs + nxt * n * (size_t)b->obs_stride; a.out.obs_stride = b->obs_stride;
    a.out.reward = b->reward + cur * n; a.out.done = b->done + cur * n;
    if (k == 0 && (rc = check_out(h, &a.out))) return rc;
    if ((rc = launch_step(h, a, s, pdl))) return rc;
    if (timed) MANSY_CUDA(cudaEventRecord(h->events[3 * k + 2], s));
  }
  return MANSY_OK;
}

int mansy_rollout_policy_host(mansy_handle_t h, mansy_policy_t p, const mansy_rollout_t *b, const mansy_rollout_host_t *host,
                              int32_t n_steps, int64_t t0, uint64_t seed, int32_t flags, void *stream) {
  if (!h || !p || !b || !host) return set_error(MANSY_E_INVALID, "NULL argument");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  if (n_steps < 0 || t0 < 0) return set_error(MANSY_E_INVALID, "n_steps / t0 must be >= 0");
  if (b->slabs < 2 || host->host_slabs < 1) return set_error(MANSY_E_INVALID, "need >= 2 device slabs and >= 1 host slab");
  if (!b->obs || !b->actions || !b->logp || !b->value || !b->reward || !b->done || !b->logits || !host->obs ||
      !host->actions || !host->logp || !host->value || !host->reward || !host->done)
    return set_error(MANSY_E_INVALID, "a rollout buffer pointer is NULL");
  if (h->dev.obs_mode == MANSY_OBS_NONE) return set_error(MANSY_E_INVALID, "the policy consumes observation rows");
  if (policy_device_of(p) != h->device) return set_error(MANSY_E_INVALID, "policy and simulator live on different devices");
  const size_t n = (size_t)h->dev.n_envs;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int is_probs = h->dev.obs_mode == MANSY_OBS_SIMPLE ? 1 : 0;
  const size_t row_bytes = (size_t)b->obs_stride * sizeof(float);
  if (!h->copy_stream) {
    MANSY_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    h->step_done.resize(kCopyRing);
    h->copy_done.resize(kCopyRing);
    for (int i = 0; i < kCopyRing; ++i) {
      MANSY_CUDA(cudaEventCreateWithFlags(&h->step_done[i], cudaEventDisableTiming));
      MANSY_CUDA(cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming));
    }
  }
  cudaStream_t cs = h->copy_stream;
  // A device slab is rewritten `slabs` steps (observations: slabs - 1) after it was produced; its copy to the host
  // must have finished by then.  The host ring is only read by the caller after this function returns.
  const int lag = b->slabs - 1 < kCopyRing ? b->slabs - 1 : kCopyRing;
  for (int32_t k = 0; k < n_steps; ++k) {
    const int64_t t = t0 + k;
    const size_t cur = (size_t)(t % b->slabs), nxt = (size_t)((t + 1) % b->slabs), hs = (size_t)(t % host->host_slabs);
    const float *obs = b->obs + cur * n * (size_t)b->obs_stride;
    int rc;
    if (k >= lag) MANSY_CUDA(cudaStreamWaitEvent(s, h->copy_done[(k - lag) % kCopyRing], 0));
    if (flags & MANSY_ROLLOUT_FP32_POLICY) {
      rc = mansy_policy_forward(p, obs, b->obs_stride, h->dev.n_envs, b->logits, b->value + cur * n, stream);
      if (!rc) rc = mansy_policy_sample(b->logits, h->dev.n_envs, is_probs, seed, t, h->dev.env_offset, b->actions + cur * n,
                                        b->logp + cur * n, stream);
    } else {
      rc = policy_forward_tc_launch(p, obs, b->obs_stride, h->dev.n_envs, b->logits, b->value + cur * n, b->actions + cur * n,
                                    b->logp + cur * n, seed, t, h->dev.env_offset, nullptr, nullptr, nullptr, 0, false, stream,
                                    getenv("MANSY_NO_POLICY_MEMO") ? nullptr : &h->dev);
    }
    if (rc) return rc;
    // act_t -> host.  The device-to-host copy engine is busy with the previous step's 12.9 MB observation slab (copy
    // stream), and a 16 KB cudaMemcpyAsync would queue behind it: ~40 us of copy-engine idle time per step once the
    // step that needs the actions has run.  When the caller's pinned buffer is mapped into the device address space
    // (cudaHostAlloc memory under unified addressing), a small kernel stores the actions straight into it instead.
    int32_t *act_mapped = nullptr;
    if (!(flags & MANSY_ROLLOUT_NO_ZERO_COPY) &&
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&act_mapped), host->actions + hs * n, 0) != cudaSuccess) {
      cudaGetLastError();
      act_mapped = nullptr;
    }
    if (act_mapped) {
      copy_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(act_mapped, b->actions + cur * n, (int32_t)n);
      count_launch();
    } else {
      MANSY_CUDA(cudaMemcpyAsync(host->actions + hs * n, b->actions + cur * n, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    MANSY_CUDA(cudaStreamSynchronize(s));                       // the host now holds act_t
    MANSY_CUDA(cudaMemcpyAsync(b->actions + cur * n, host->actions + hs * n, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.actions = b->actions + cur * n; a.n = h->dev.n_envs; a.auto_reset = 1; a.n_steps = 1;
    a.out.obs = b->obs + nxt * n * (size_t)b->obs_stride; a.out.obs_stride = b->obs_stride;
    a.out.reward = b->reward + cur * n; a.out.done = b->done + cur * n;
    if (k == 0 && (rc = check_out(h, &a.out))) return rc;
    if ((rc = launch_step(h, a, s))) return rc;
    MANSY_CUDA(cudaEventRecord(h->step_done[k % kCopyRing], s));
    MANSY_CUDA(cudaStreamWaitEvent(cs, h->step_done[k % kCopyRing], 0));
    float *reward_m = nullptr, *logp_m = nullptr, *value_m = nullptr;
    uint8_t *done_m = nullptr;
    const bool scalars_mapped = !(flags & MANSY_ROLLOUT_NO_ZERO_COPY) &&
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&reward_m), host->reward + hs * n, 0) == cudaSuccess &&
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&done_m), host->done + hs * n, 0) == cudaSuccess &&
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&logp_m), host->logp + hs * n, 0) == cudaSuccess &&
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&value_m), host->value + hs * n, 0) == cudaSuccess;
    if (scalars_mapped) {                                        // after the event: the slab copy does not wait for it
      store_step_scalars_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reward_m, done_m, logp_m, value_m, a.out.reward, a.out.done,
                                                                            b->logp + cur * n, b->value + cur * n, (int32_t)n);
      count_launch();
    } else {
      cudaGetLastError();
    }
    MANSY_CUDA(cudaMemcpyAsync(host->obs + hs * n * (size_t)b->obs_stride, a.out.obs, n * row_bytes, cudaMemcpyDeviceToHost, cs));
    if (!scalars_mapped) {
      MANSY_CUDA(cudaMemcpyAsync(host->reward + hs * n, a.out.reward, n * sizeof(float), cudaMemcpyDeviceToHost, cs));
      MANSY_CUDA(cudaMemcpyAsync(host->done + hs * n, a.out.done, n, cudaMemcpyDeviceToHost, cs));
      MANSY_CUDA(cudaMemcpyAsync(host->logp + hs * n, b->logp + cur * n, n * sizeof(float), cudaMemcpyDeviceToHost, cs));
      MANSY_CUDA(cudaMemcpyAsync(host->value + hs * n, b->value + cur * n, n * sizeof(float), cudaMemcpyDeviceToHost, cs));
    }
    MANSY_CUDA(cudaEventRecord(h->copy_done[k % kCopyRing], cs));
  }
  MANSY_CUDA(cudaStreamSynchronize(cs));                        // every result of the n_steps is in host memory
  MANSY_CUDA(cudaStreamSynchronize(s));
  return MANSY_OK;
}

int mansy_rollout_kernel_ms(mansy_handle_t h, double *policy_ms, double *step_ms, int32_t *n_steps) {
  if (!h || !policy_ms || !step_ms || !n_steps) return set_error(MANSY_E_INVALID, "NULL argument");
  double pm = 0.0, sm = 0.0;
  for (int k = 0; k < h->timed_steps; ++k) {
    float a = 0.f, c = 0.f;
    MANSY_CUDA(cudaEventElapsedTime(&a, h->events[3 * k], h->events[3 * k + 1]));
    MANSY_CUDA(cudaEventElapsedTime(&c, h->events[3 * k + 1], h->events[3 * k + 2]));
    pm += a; sm += c;
  }
  *policy_ms = pm; *step_ms = sm; *n_steps = h->timed_steps;
  return MANSY_OK;
}

int mansy_episode_stats(mansy_handle_t h, double *stats_dev, void *stream) {
  if (!h || !stats_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  MANSY_CUDA(cudaMemcpyAsync(stats_dev, h->dev.stats, (size_t)h->dev.n_envs * MANSY_STATS_DOUBLES * sizeof(double),
                             cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return MANSY_OK;
}

int mansy_set_outcome_table(mansy_handle_t h, int32_t enable) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  if (enable && !h->outcome) return set_error(MANSY_E_STATE, "this handle has no outcome table (too large, or MANSY_NO_OUTCOME_TABLE=1)");
  h->dev.outcome = enable ? h->outcome : nullptr;
  return MANSY_OK;
}

int mansy_stats_clear(mansy_handle_t h, void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "handle is NULL");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  const size_t n = (size_t)h->dev.n_envs * MANSY_STATS_DOUBLES;
  stats_clear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(h->dev.stats, n);
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

int mansy_expert_actions(mansy_handle_t h, int32_t horizon, int32_t *actions_dev, double *best_value_dev, void *stream) {
  if (!h || !actions_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  if (horizon < 1 || horizon > kExpertMaxHorizon) return set_error(MANSY_E_INVALID, "horizon must be 1..6");
  expert_mpc_kernel<<<(unsigned)h->dev.n_envs, kExpertThreads, 0, static_cast<cudaStream_t>(stream)>>>(h->dev, horizon, actions_dev,
                                                                                                       best_value_dev);
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

int mansy_state_snapshot(mansy_handle_t h, void *state_dev, void *stream) {
  if (!h || !state_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  MANSY_CUDA(cudaMemcpyAsync(state_dev, h->dev.state, (size_t)h->dev.n_envs * sizeof(EnvState),
                             cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return MANSY_OK;
}

int mansy_error_flag(mansy_handle_t h, int32_t *flag_host) {
  if (!h || !flag_host) return set_error(MANSY_E_INVALID, "NULL argument");
  DeviceScope dscope(h->device);
  MANSY_CUDA(dscope.err);
  MANSY_CUDA(cudaMemcpy(flag_host, h->dev.error_flag, sizeof(int32_t), cudaMemcpyDeviceToHost));
  return MANSY_OK;
}

int mansy_viewport_tiles(const float *gt_xy_dev, const float *pred_xy_dev, int64_t n_chunks, int32_t points,
                         int32_t video_width, int32_t video_height, int32_t fov_width, int32_t fov_height,
                         uint64_t *gt_mask_dev, uint64_t *pred_mask_dev, double *acc_dev, void *stream) {
  if (!gt_xy_dev || !gt_mask_dev) return set_error(MANSY_E_INVALID, "gt pointers are NULL");
  DeviceScope dscope(device_of_pointer(gt_xy_dev));
  if ((pred_xy_dev != nullptr) != (pred_mask_dev != nullptr) || (pred_xy_dev != nullptr) != (acc_dev != nullptr))
    return set_error(MANSY_E_INVALID, "pred_xy, pred_mask and acc must be given together");
  if (n_chunks < 0 || points < 1) return set_error(MANSY_E_INVALID, "bad n_chunks / points");
  if (video_width < 8 || video_height < 8 || video_width % 8 || video_height % 8)
    return set_error(MANSY_E_INVALID, "frame must split into 8x8 tiles");
  if (fov_width < 0 || fov_height < 0 || fov_width > video_width || fov_height > video_height)
    return set_error(MANSY_E_INVALID, "FoV larger than the frame is outside the reference's nine cases");
  if (n_chunks == 0) return MANSY_OK;
  const int threads = 256;
  const int64_t grid = (n_chunks + threads - 1) / threads;
  viewport_tiles_kernel<<<(unsigned)grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      gt_xy_dev, pred_xy_dev, n_chunks, points, video_width, video_height, fov_width, fov_height, gt_mask_dev,
      pred_mask_dev, acc_dev, nullptr);
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

int mansy_allocate_tile_versions(const uint64_t *masks_dev, const int32_t *actions_dev, int64_t n,
                                 const int32_t video_rates[5], uint8_t *versions_dev, void *stream) {
  if (!masks_dev || !actions_dev || !versions_dev || !video_rates) return set_error(MANSY_E_INVALID, "NULL argument");
  if (n < 0) return set_error(MANSY_E_INVALID, "n < 0");
  if (n == 0) return MANSY_OK;
  DeviceScope dscope(device_of_pointer(masks_dev));
  uint32_t lut[5];
  build_rate_lut(video_rates, lut);
  const int threads = 128;
  allocate_versions_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      masks_dev, actions_dev, n, lut[0], lut[1], lut[2], lut[3], lut[4], versions_dev);
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

// ---- CPU self-tests of the shared __host__ __device__ building blocks (no GPU needed) --------
// These run the same mansy_core.cuh code the kernels run, on the host, so the non-GPU test-suite
// can compare it with the oracle.  They are not a product path.
int mansy_selftest_allocate(uint64_t mask, int32_t action, const int32_t video_rates[5], uint8_t versions_out[64]) {
  uint32_t lut[5];
  build_rate_lut(video_rates, lut);
  int rin, rout;
  action_to_rates(action, rin, rout);
  const TileScaleMasks dm = tile_scale_masks(mask);
  for (int t = 0; t < 64; ++t) versions_out[t] = (uint8_t)tile_version(lut[rout], rin, tile_scale(dm, t));
  return MANSY_OK;
}

int mansy_selftest_fov_mask(int32_t x, int32_t y, int32_t w, int32_t hgt, int32_t fov_w, int32_t fov_h,
                            uint64_t *mask_out, int32_t *valid_out) {
  bool valid = true;
  *mask_out = fov_tile_mask(x, y, w, hgt, fov_w, fov_h, valid);
  *valid_out = valid ? 1 : 0;
  return MANSY_OK;
}

int mansy_selftest_centre_to_pixel(float v, int32_t length) { return centre_to_pixel(v, length); }

int mansy_selftest_download(const double *thr, int32_t trace_len, int64_t size, int32_t *cur_idx, double *cur_time,
                            double *buf, double *download_time, double *rebuffer) {
  bool ok = true;
  int idx = *cur_idx;
  double tm = *cur_time;
  const double dl = trace_download((double)size, TracePtr{thr}, trace_len, idx, tm, ok);
  *cur_idx = idx; *cur_time = tm;
  *download_time = dl;
  *rebuffer = buffer_push(*buf, 1.0, dl);
  return ok ? MANSY_OK : MANSY_E_STATE;
}

int mansy_selftest_ddiv_rcp(const double *a, const double *b, int64_t n, double *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = ddiv_rcp(a[i], b[i], 1.0 / b[i]);
  return MANSY_OK;
}

int mansy_selftest_hashed_action(uint64_t seed, uint64_t env, uint64_t step) { return hashed_action(seed, env, step, kActions); }

}  // extern "C"
