// mansy_policy.cuh -- shared between the fp32 CUDA-core policy kernel (mansy_policy.cu) and the
// tcgen05 TF32 kernel (mansy_policy_tc.cu).
#pragma once
#include <string>
#include <vector>

#include "mansy_sim.cuh"

namespace mansy {

int set_error(int code, const std::string &msg);
void count_launch();
int sim_device_of(mansy_handle_t h);     // mansy_sim.cu

constexpr int kHidden = 128;
constexpr int kMaxBranches = 10;
constexpr float kLeaky = 0.01f;       // torch.nn.LeakyReLU default negative_slope

struct PolicyDev {
  int32_t kind;
  int32_t n_branches;
  int32_t feat_dim;                    // 128 * n_branches
  int32_t residual_branch;             // MANSY: 9 (qoe_weight features), SimpleRL: -1
  int32_t softmax;                     // SimpleRL actor returns probabilities
  int32_t obs_off[kMaxBranches];       // offset of the branch's segment in the observation row
  int32_t k[kMaxBranches];             // segment length
  int32_t w_off[kMaxBranches];         // offset of the branch's [K][128] block in w1t
  const float *w1t;                    // layer-1 weights, K-major per branch
  const float *b1;                     // [n_branches][128]
  const float *wfct;                   // [feat_dim][256]: columns 0..127 actor.fc, 128..255 critic.fc
  const float *bfc;                    // [256]
  const float *wout;                   // [16][128]: rows 0..14 actor.out, row 15 critic.out
  const float *bout;                   // [16]
};

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : kLeaky * x; }

// Categorical(logits).sample() (bitrate_selection/run_mansy.py:228-229): inverse CDF on the
// softmax with a uniform from a counter-based hash keyed by (seed, global env, step).
__device__ __forceinline__ float categorical_uniform(uint64_t seed, uint64_t env, uint64_t step) {   // in [0, 1)
  uint64_t z = seed * 0x9E3779B97F4A7C15ULL + env * 0xBF58476D1CE4E5B9ULL + step * 0x94D049BB133111EBULL +
               0x2545F4914F6CDD1DULL;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27; z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
// Second half of the sample: `p` holds unnormalised weights, `s` their sum (added in index order), `u01` the uniform.
__device__ __forceinline__ void categorical_pick(const float (&p)[kActions], float s, float u01, int &act, float &logp) {
  const float u = u01 * s;                                         // uniform in [0, s)
  act = kActions - 1;
  float cum = 0.f;
  bool found = false;
#pragma unroll
  for (int o = 0; o < kActions; ++o) {
    cum += p[o];
    if (!found && u < cum) { act = o; found = true; }
  }
  float pa = p[0];
#pragma unroll
  for (int o = 1; o < kActions; ++o) if (o == act) pa = p[o];
  logp = logf(pa / s);
}
// `p` holds logits (is_probs == 0) or probabilities; on return it holds unnormalised weights.
__device__ __forceinline__ void categorical_sample(float (&p)[kActions], int is_probs, uint64_t seed, uint64_t env,
                                                   uint64_t step, int &act, float &logp) {
  float s = 0.f;
  if (!is_probs) {
    float m = p[0];
#pragma unroll
    for (int o = 1; o < kActions; ++o) m = fmaxf(m, p[o]);
#pragma unroll
    for (int o = 0; o < kActions; ++o) { p[o] = expf(p[o] - m); s += p[o]; }
  } else {
#pragma unroll
    for (int o = 0; o < kActions; ++o) s += p[o];
  }
  categorical_pick(p, s, categorical_uniform(seed, env, step), act, logp);
}

struct TcState;   // tensor-core path state (mansy_policy_tc.cu)

}  // namespace mansy

struct mansy_policy {
  mansy::PolicyDev dev;
  int device = 0;
  size_t smem_bytes = 0;
  std::vector<void *> allocs;
  mansy::TcState *tc = nullptr;
  // Memoised table branches (policy_memo_for): the 320-input branches (next chunk sizes / qualities) see rows of the
  // simulator's read-only tables, so their contribution to the hidden pre-activation is a function of (video, chunk)
  float *memo = nullptr;            // [memo_rows][256]
  uint64_t memo_key = 0;            // SimDev.uid of the simulator the table was built for
  int memo_rows = 0, memo_cap = 0;
};

namespace mansy {
// Builds the tensor-core weight images / tensor maps for `p`; returns MANSY_OK or an error code
// (the fp32 path keeps working when this fails; mansy_policy_forward_tc then reports the reason).
int tc_create(mansy_policy *p, const mansy_policy_weights_t *w);
void tc_destroy(mansy_policy *p);
// Tensor-core forward launch shared by the C entry points and the rollout loops (mansy_policy_tc.cu).
// `memo_sim` != NULL: the rows are the CURRENT observations of that simulator's environments (row i = env i); the
// 320-input table branches then come from the (video, chunk) memo instead of the tensor pipe (policy_memo_for).
int policy_forward_tc_launch(mansy_policy *p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                             float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                             int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, int64_t *timeline_dev,
                             int32_t timeline_cta, bool pdl, void *stream, const SimDev *memo_sim = nullptr);
// Device table memo[video * n_chunks + chunk][256] for (policy, simulator tables): sum over the 320-input branches of
// LeakyReLU(W1_b x_b + b1_b) . [actor.fc | critic.fc]_b^T in exact fp32, built on `stream` the first time the pair
// is seen (and again when the policy meets another simulator).  *memo_out = NULL when the net has no such branch
// input in the tables (the QoE identifier is never evaluated inside a rollout).
int policy_memo_for(mansy_policy *p, const SimDev &S, void *stream, const float **memo_out);
// n_steps x (policy + sample + simulator step) in one launch of the cluster kernel; *launched = 0 if not applicable.
int rollout_fused_launch(mansy_policy *p, const SimDev &S, const mansy_rollout_t *b, int32_t n_steps, int64_t t0,
                         uint64_t seed, void *stream, int *launched);
}  // namespace mansy
