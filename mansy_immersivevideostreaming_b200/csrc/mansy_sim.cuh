// mansy_sim.cuh -- device-side data layout of the simulator (shared by the .cu files).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include "mansy_b200.h"
#include "mansy_core.cuh"

namespace mansy {

constexpr int kTiles = 64;
constexpr int kRates = 5;
constexpr int kPastK = 8;
constexpr int kActions = 15;
constexpr int kLanesPerEnv = 8;    // one lane per row of the 8x8 tile grid
constexpr int kThreadsPerBlock = 128;
constexpr int kEnvsPerBlock = kThreadsPerBlock / kLanesPerEnv;
constexpr int kTableRow = kRates * kTiles;  // 320 values per (video, chunk)
constexpr int kNoAction = 255;
constexpr int kTraceWindow = 8;    // trace entries an environment's 8 lanes prefetch per load (device rows are wrap-padded by this)

// Per-environment state: one 128-byte line, read with 8 broadcast 128-bit loads by the 8 lanes of
// the env's group and written back one quad per lane.
struct __align__(16) EnvState {
  double cur_time;      // q0  network.py cur_time
  double buf;           //     buffer.py buf_size
  double prev_vq;       // q1  qoe.py prev_viewport_quality (valid when ep_step > 0)
  int32_t next_chunk;   //     simulator.py next_chunk
  int32_t cur_idx;      //     network.py cur_idx
  int32_t ep_step;      // q2  steps taken in this episode
  int32_t cursor;       //     mansy_env.py worker_id (next sample to use)
  int32_t video;        //     table indices of the episode's sample
  int32_t pair;         //     video * n_users + user
  int32_t trace;        // q3
  int32_t end_chunk;    //     min(last viewport chunk, Video_Time - 1)  (simulator.py:41-42)
  int32_t sample_id;
  int32_t flags;        //     bit 0: episode finished (awaiting reset); bits 8..15: last action (255 = none)
  float w0, w1, w2;     // q4  qoe weights of the sample
  int32_t start_chunk;  //     first chunk of the viewport list (hmdtrace.py:10)
  double sum_qoe;       // q5  running per-episode sums (mansy_env.py:187-190 log lists)
  double sum_q1;
  double sum_q2;        // q6
  double sum_q3;
  double ep_return;     // q7  running sum of rewards
  double reserved;
};
static_assert(sizeof(EnvState) == 128, "EnvState must be one 128-byte line");

constexpr int kFlagFinished = 1;

// History ring: hist[env][slot][8]; the value pushed at episode step s lives in slot s & 7.
// Slot layout (floats): throughput, rate_in, rate_out, accuracy, viewport quality (qoe1),
// quality variance (qoe3), rebuffer/startup (qoe2/5), raw rebuffer (qoe2; SimpleRL observation).
constexpr int kHistFloatsPerEnv = 64;

// What the download / QoE part of a step needs from (viewport pair, chunk, action): a pure function of the read-only
// tables (utils/common.py:101-193 allocation, simulator.py:94-101 gather + sums, qoe.py:23-28), tabulated once per
// handle by the same device code a step runs (outcome_build_kernel) -- the reference's own ExpertEnv caches the same
// per-(chunk, action) statistics (expert_env.py:121-160).  16 actions: 0..14 and the out-of-table action (rates 0, 0).
struct __align__(16) ChunkOutcome {
  double q1;       // viewport quality / max_quality                     (qoe.py:23,26)
  double intra;    // sum |q - vq| over the actual viewport / count / max_quality   (qoe.py:24-25)
  int32_t size;    // chunk bytes: exact integer sum of the 64 selected tiles       (simulator.py:100)
  int32_t pad0;
  int64_t pad1;
};
static_assert(sizeof(ChunkOutcome) == 32, "ChunkOutcome is two 16-byte quads");
constexpr int kOutcomeActions = 16;

// Everything reset_episode needs about a sample (mansy_env.py:99-134, simulator.py:15-46), resolved once per handle from
// the sample list and the tables: one 48-byte record instead of a chain of dependent lookups at every episode end.
struct __align__(16) EpisodeInit {
  int32_t video, pair, trace, end_chunk;      // end_chunk = min(last viewport chunk, Video_Time - 1)   (simulator.py:41-42)
  float w0, w1, w2;                           // QoE weights of the sample
  int32_t start_chunk;                        // first chunk of the viewport list (hmdtrace.py:10)
  uint64_t first_pred;                        // predicted viewport of the first observation's chunk (mansy_env.py:119-121)
  uint64_t pad;
};
static_assert(sizeof(EpisodeInit) == 48, "EpisodeInit is three 16-byte quads");

struct SimDev {
  // read-only tables
  const int32_t *size;
  const float *quality;
  const float *size_norm;   // float(size) / float(max_size)           (utils/common.py:45-47)
  const float *qual_norm;   // quality / float(video_rates[-1])        (utils/common.py:40-42)
  const float *obs_tab;     // [videos][chunks][640]: size_norm row | qual_norm row -- columns 8..647 of a MANSY observation, one bulk copy
  const int32_t *video_time;
  const uint64_t *vp_gt;
  const uint64_t *vp_pred;
  const uint32_t *vp_scale;  // [pairs][n_vp_chunks][8]: per tile row, 8 x 4-bit pyramid scales derived from vp_pred (step_env)
  const double *vp_acc;
  const int32_t *vp_start;
  const int32_t *vp_end;
  const double *trace;
  const double *trace_rcp;   // RN(1 / trace[i]) for ddiv_rcp (same layout as trace)
  const ChunkOutcome *outcome;  // [pairs][n_vp_chunks][16] or NULL (table disabled / too large): step_env then gathers
  const int32_t *trace_len;
  const float *qoe_w;
  const int32_t *samples;
  const EpisodeInit *ep_init;   // [n_samples]
  int32_t n_videos, n_chunks, n_users, n_vp_chunks, n_traces, trace_stride, n_qoe, n_samples;
  // per-env state
  EnvState *state;
  float *hist;
  double *stats;            // [n_envs][MANSY_STATS_DOUBLES]
  int32_t *error_flag;      // set by kernels on data errors (trace that can never finish a download)
  uint64_t uid;             // unique per handle in this process (keys caches derived from the tables, e.g. the policy memo)
  // configuration
  int32_t n_envs, env_offset, worker_num, obs_mode, reward_mode;
  int32_t startup_download;
  double chunk_length, max_quality, max_throughput, startup_d;
  double rcp_max_quality, rcp_max_throughput, rcp_startup_d;   // correctly rounded reciprocals (ddiv_rcp)
  double rcp_count[kTiles + 1];                                // RN(1 / k), k = popcount of the actual viewport
  float startup_f;
  float rate_norm_hist[kRates];  // float(rates[r] / rates[-1]) with a float64 division (mansy_env.py:197,199)
  float rate_norm_f32[kRates];   // float(rates[r]) / float(rates[-1])          (simple_rl_env.py:137-138)
  uint32_t lut[kRates];
};

// Entry points without a handle run where their (device) buffers live.
inline int device_of_pointer(const void *dev_ptr) {
  cudaPointerAttributes a;
  int cur = 0;
  if (cudaPointerGetAttributes(&a, dev_ptr) == cudaSuccess && a.type == cudaMemoryTypeDevice) return a.device;
  cudaGetLastError();
  cudaGetDevice(&cur);
  return cur;
}

// Entry points run on their handle's device whatever the caller's current device is (a process driving several GPUs) and
// leave the caller's current device as they found it.
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceScope(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      switched = err == cudaSuccess;
    }
  }
  ~DeviceScope() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope &) = delete;
  DeviceScope &operator=(const DeviceScope &) = delete;
};

}  // namespace mansy
