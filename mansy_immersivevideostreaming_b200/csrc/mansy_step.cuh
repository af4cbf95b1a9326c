// mansy_step.cuh -- device code of one simulator chunk-step (shared by the stand-alone step / reset kernels in
// mansy_sim.cu and the fused policy + step rollout kernel in mansy_policy_tc.cu).
//
// Reference map (file:line relative to the reference root):
//   step_env        bitrate_selection/envs/mansy_env.py:154-248, envs/simple_rl_env.py:113-160,
//                   simulators/simulator.py:88-108, simulators/network.py:22-35,
//                   simulators/buffer.py:8-15, utils/qoe.py:22-34, utils/common.py:101-193
//   reset_episode   envs/mansy_env.py:99-152, simulators/simulator.py:15-46
//   emit_obs        envs/mansy_env.py:136-150,232-246, envs/simple_rl_env.py:103-109,152-158
#pragma once
#include "mansy_sim.cuh"

namespace mansy {

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
struct StepArgs {
  const int32_t *actions;   // [n] (action_mode 0)
  const int32_t *env_ids;   // [n] or NULL
  int32_t n;
  int32_t auto_reset;
  int32_t action_mode;      // 0: actions array, 1: hashed_action(seed, global env, step)
  int32_t n_steps;          // steps executed inside one launch (state stays in registers)
  uint64_t seed;
  int64_t step0;
  int64_t rows_per_step;    // output rows between consecutive steps (0: overwrite the same rows)
  mansy_out_t out;
};

__device__ __forceinline__ unsigned group_mask() { return 0xFFu << ((threadIdx.x & 31) & ~7); }

__device__ __forceinline__ int group_sum(int v, unsigned m) {
  v += __shfl_xor_sync(m, v, 1);
  v += __shfl_xor_sync(m, v, 2);
  v += __shfl_xor_sync(m, v, 4);
  return v;
}
__device__ __forceinline__ double group_sum(double v, unsigned m) {
  v = dadd(v, __shfl_xor_sync(m, v, 1));
  v = dadd(v, __shfl_xor_sync(m, v, 2));
  v = dadd(v, __shfl_xor_sync(m, v, 4));
  return v;
}

union StateQuads {
  EnvState s;
  uint4 q[8];
  __device__ StateQuads() {}
};

__device__ __forceinline__ void load_state(const SimDev &S, int e, EnvState &st) {
  StateQuads u;
  const uint4 *p = reinterpret_cast<const uint4 *>(S.state + e);
#pragma unroll
  for (int i = 0; i < 8; ++i) u.q[i] = p[i];   // same address in all 8 lanes: broadcast
  st = u.s;
}

__device__ __forceinline__ void store_state(const SimDev &S, int e, const EnvState &st, int sub) {
  StateQuads u;
  u.s = st;
  uint4 v = u.q[0];
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (sub == i) v = u.q[i];
  reinterpret_cast<uint4 *>(S.state + e)[sub] = v;   // lane i writes quad i: one 128-byte line per env
}

__device__ __forceinline__ void load_slot(const SimDev &S, int e, int sub, float (&slot)[8]) {
  const float4 *p = reinterpret_cast<const float4 *>(S.hist + (size_t)e * kHistFloatsPerEnv + sub * 8);
  const float4 a = p[0], b = p[1];
  slot[0] = a.x; slot[1] = a.y; slot[2] = a.z; slot[3] = a.w;
  slot[4] = b.x; slot[5] = b.y; slot[6] = b.z; slot[7] = b.w;
}

__device__ __forceinline__ void store_slot(const SimDev &S, int e, int sub, const float (&slot)[8]) {
  float4 *p = reinterpret_cast<float4 *>(S.hist + (size_t)e * kHistFloatsPerEnv + sub * 8);
  p[0] = make_float4(slot[0], slot[1], slot[2], slot[3]);
  p[1] = make_float4(slot[4], slot[5], slot[6], slot[7]);
}

// envs/mansy_env.py:99-134 + simulators/simulator.py:15-46: pick the next sample and rebuild the
// episode state.  Executed identically by the 8 lanes of the group.  `ei` = S.ep_init[st.cursor % S.n_samples].
__device__ __forceinline__ void reset_episode(const SimDev &S, EnvState &st, const EpisodeInit &ei) {
  st.sample_id = st.cursor % S.n_samples;
  st.cursor = (st.cursor + S.worker_num) % S.n_samples;          // mansy_env.py:100-101
  st.video = ei.video;
  st.pair = ei.pair;
  st.trace = ei.trace;
  st.w0 = ei.w0; st.w1 = ei.w1; st.w2 = ei.w2;
  st.buf = S.chunk_length * 3.0;                                  // buffer.py:6
  st.cur_time = 0.0;                                              // network.py:19-20
  st.cur_idx = 0;
  st.next_chunk = S.startup_download + 1;                         // simulator.py:45
  st.start_chunk = ei.start_chunk;
  st.end_chunk = ei.end_chunk;                                    // simulator.py:41-42
  st.prev_vq = 0.0;
  st.ep_step = 0;
  st.flags = kNoAction << 8;
  st.sum_qoe = st.sum_q1 = st.sum_q2 = st.sum_q3 = 0.0;
  st.ep_return = 0.0;
}
__device__ __forceinline__ EpisodeInit load_episode_init(const SimDev &S, int cursor) {
  const uint4 *p = reinterpret_cast<const uint4 *>(S.ep_init + cursor % S.n_samples);
  union { EpisodeInit e; uint4 q[3]; } u;
  u.q[0] = __ldg(p); u.q[1] = __ldg(p + 1); u.q[2] = __ldg(p + 2);
  return u.e;
}
__device__ __forceinline__ void reset_episode(const SimDev &S, EnvState &st) { reset_episode(S, st, load_episode_init(S, st.cursor)); }

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ double f2_as_double(float lo, float hi) {
  return __hiloint2double((int)__float_as_uint(hi), (int)__float_as_uint(lo));
}

// network.py:22-35 with the trace window held by the environment's 8 lanes: lane j of the group owns entry
// cur_idx + j of the wrap-padded row and its reciprocal (one load each per lane, issued by the caller before the
// dependent work), the sequential walk reads entry `pos` with a shuffle and only reloads every 8 seconds of simulated
// download.  Arithmetic and operation order are those of trace_download (mansy_core.cuh) -- the one division of a
// download, size / thr in the last, partial second, is ddiv_rcp on the tabulated reciprocal (bit-identical to an
// IEEE division) -- so results are bit-identical to the scalar loop.
__device__ __forceinline__ double trace_download_window(double size, double win, double rwin, const double *__restrict__ tr,
                                                        const double *__restrict__ trr, int trace_len, int &cur_idx,
                                                        double &cur_time, int sub, unsigned gmask, bool &ok) {
  const double start = cur_time;
  const int base = (threadIdx.x & 31) & ~7;
  int pos = 0, it = 0;
  if (gmask == 0xFFFFFFFFu) {
    // whole warp converged (the mask is a compile-time constant here): one warp-uniform loop over the longest of
    // the four downloads, plain full-mask shuffles instead of the partial-mask collective sequence
    bool more = size > 0.0;
    while (true) {
      if (!__any_sync(0xFFFFFFFFu, more)) break;
      if (pos == kTraceWindow) {
        win = __ldg(tr + cur_idx + sub);
        rwin = __ldg(trr + cur_idx + sub);
        pos = 0;
      }
      const double thr = __shfl_sync(0xFFFFFFFFu, win, base + pos);
      if (more) {
        const double next_tick = floor(dadd(cur_time, 1.0));
        const double remain = dmul(dsub(next_tick, cur_time), thr);
        if (size >= remain) {
          cur_idx = (cur_idx + 1 == trace_len) ? 0 : cur_idx + 1;
          cur_time = next_tick;
          size = dsub(size, remain);
          ++pos;
          if (++it > (1 << 22)) { ok = false; more = false; }
          else more = size > 0.0;
        } else {
          more = false;          // the last, partial second: one division, after the loop (`pos` stays on its entry)
        }
      }
    }
    const double thr = __shfl_sync(0xFFFFFFFFu, win, base + pos);
    const double rthr = __shfl_sync(0xFFFFFFFFu, rwin, base + pos);
    if (size > 0.0 && ok) cur_time = dadd(cur_time, ddiv_rcp(size, thr, rthr));
    return dsub(cur_time, start);
  }
  while (size > 0.0) {
    if (pos == kTraceWindow) {
      win = __ldg(tr + cur_idx + sub);
      rwin = __ldg(trr + cur_idx + sub);
      pos = 0;
    }
    const double thr = __shfl_sync(gmask, win, base + pos);
    const double next_tick = floor(dadd(cur_time, 1.0));
    const double remain = dmul(dsub(next_tick, cur_time), thr);
    if (size >= remain) {
      cur_idx = (cur_idx + 1 == trace_len) ? 0 : cur_idx + 1;
      cur_time = next_tick;
      size = dsub(size, remain);
      ++pos;
    } else {
      const double rthr = __shfl_sync(gmask, rwin, base + pos);
      cur_time = dadd(cur_time, ddiv_rcp(size, thr, rthr));
      size = 0.0;
    }
    if (++it > (1 << 22)) { ok = false; break; }
  }
  return dsub(cur_time, start);
}

// Observation row from the env state (a pure function of state + history ring), in two parts: the rows copied
// from the read-only tables (next chunk sizes / qualities, predicted viewport), which depend only on WHICH chunk
// comes next, and the dynamic part (history columns, last action, buffer, qoe weights).  The fused rollout
// kernel writes the table part early (the next chunk is known before the action is), everything else calls
// emit_obs.
// Predicted-viewport columns of the row (mansy_env.py:229,236 / simple_rl_env.py:156): bit t of the mask -> float t.
template <int MODE>
__device__ __forceinline__ void emit_obs_pred(uint64_t pred, int sub, float *__restrict__ row) {
  const uint32_t pbyte = (uint32_t)(pred >> (8 * sub)) & 0xFFu;
  const float4 p0 = make_float4((float)(pbyte & 1u), (float)((pbyte >> 1) & 1u), (float)((pbyte >> 2) & 1u),
                                (float)((pbyte >> 3) & 1u));
  const float4 p1 = make_float4((float)((pbyte >> 4) & 1u), (float)((pbyte >> 5) & 1u),
                                (float)((pbyte >> 6) & 1u), (float)((pbyte >> 7) & 1u));
  float4 *dp = reinterpret_cast<float4 *>(row + (MODE == MANSY_OBS_MANSY ? 648 : 328));
  dp[2 * sub] = p0;
  dp[2 * sub + 1] = p1;
}

template <int MODE>
__device__ __forceinline__ void emit_obs_tables(const SimDev &S, int video, int pair, int start_chunk, int obs_chunk, int sub,
                                                float *__restrict__ row) {
  const uint64_t pred = __ldg(S.vp_pred + (size_t)pair * S.n_vp_chunks + (obs_chunk - start_chunk));
  const size_t tab = ((size_t)video * S.n_chunks + obs_chunk) * kTableRow;
  const float4 *s4 = reinterpret_cast<const float4 *>(S.size_norm + tab);
  float4 *ds = reinterpret_cast<float4 *>(row + 8);
  float4 tv[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) tv[i] = __ldg(s4 + sub + 8 * i);
  if (MODE == MANSY_OBS_MANSY) {
    const float4 *q4 = reinterpret_cast<const float4 *>(S.qual_norm + tab);
    float4 *dq = reinterpret_cast<float4 *>(row + 328);
    float4 tq[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) tq[i] = __ldg(q4 + sub + 8 * i);
#pragma unroll
    for (int i = 0; i < 10; ++i) ds[sub + 8 * i] = tv[i];
#pragma unroll
    for (int i = 0; i < 10; ++i) dq[sub + 8 * i] = tq[i];
  } else {  // MANSY_OBS_SIMPLE
#pragma unroll
    for (int i = 0; i < 10; ++i) ds[sub + 8 * i] = tv[i];
  }
  emit_obs_pred<MODE>(pred, sub, row);
}

// `wn`: the normalised QoE weights w / (w0 + w1 + w2) when the caller has them already (they only change with the sample).
template <int MODE>
__device__ __forceinline__ void emit_obs_dynamic(const SimDev &S, const EnvState &st, const float (&slot)[8], int sub,
                                                 unsigned gmask, float *__restrict__ row, const float *wn = nullptr) {
  const int pushes = st.ep_step;
  const int newest = (pushes - 1) & 7;
  const int k = (newest - sub) & 7;         // observation index of this lane's slot (0 = newest)
  const bool valid = k < pushes;            // older entries are still the zeros of reset
  const int la = (st.flags >> 8) & 0xFF;    // last action: 0..14, 15 = out-of-table action, 255 = none
  if (MODE == MANSY_OBS_MANSY) {
    row[0 + k] = valid ? slot[0] : 0.f;      // throughput
    row[712 + k] = valid ? slot[1] : 0.f;    // rates_inside
    row[720 + k] = valid ? slot[2] : 0.f;    // rates_outside
    row[728 + k] = valid ? slot[3] : 0.f;    // viewport_acc
    row[736 + k] = valid ? slot[4] : 0.f;    // past_viewport_qualities
    row[744 + k] = valid ? slot[5] : 0.f;    // past_quality_variances
    row[752 + k] = valid ? slot[6] : 0.f;    // past_rebuffering
    if (sub < 4) {            // action_one_hot (15 + 1 pad)
      const int b = 4 * sub;
      reinterpret_cast<float4 *>(row + 760)[sub] =
          make_float4(la == b ? 1.f : 0.f, la == b + 1 ? 1.f : 0.f, la == b + 2 ? 1.f : 0.f,
                      (la == b + 3 && b + 3 < kActions) ? 1.f : 0.f);
    } else if (sub == 4) {    // qoe_weight (utils/common.py:55-57) and buffer / startup_download
      if (wn) {
        reinterpret_cast<float4 *>(row + 776)[0] = make_float4(wn[0], wn[1], wn[2], fdiv((float)st.buf, S.startup_f));
      } else {
        const float ws = (float)dadd(dadd((double)st.w0, (double)st.w1), (double)st.w2);
        reinterpret_cast<float4 *>(row + 776)[0] =
            make_float4(fdiv(st.w0, ws), fdiv(st.w1, ws), fdiv(st.w2, ws), fdiv((float)st.buf, S.startup_f));
      }
    } else if (sub == 5) {
      reinterpret_cast<float4 *>(row + 780)[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {  // MANSY_OBS_SIMPLE
    const float rb_newest = __shfl_sync(gmask, slot[7], ((threadIdx.x & 31) & ~7) + newest);
    row[0 + k] = valid ? slot[0] : 0.f;
    if (sub == 0) {           // last_bitrates (2), rebuffer (1), pad
      float lb0 = 0.f, lb1 = 0.f;
      if (la != kNoAction) {
        int rin, rout;
        action_to_rates(la, rin, rout);
        lb0 = S.rate_norm_f32[rin];
        lb1 = S.rate_norm_f32[rout];
      }
      reinterpret_cast<float4 *>(row + 392)[0] = make_float4(lb0, lb1, pushes > 0 ? rb_newest : 0.f, 0.f);
    } else if (sub == 1) {
      reinterpret_cast<float4 *>(row + 396)[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int MODE>
__device__ __forceinline__ void emit_obs(const SimDev &S, const EnvState &st, const float (&slot)[8], int sub,
                                         unsigned gmask, float *__restrict__ row) {
  emit_obs_tables<MODE>(S, st.video, st.pair, st.start_chunk, min(st.next_chunk, st.end_chunk), sub, row);  // terminal obs repeats the last chunk
  emit_obs_dynamic<MODE>(S, st, slot, sub, gmask, row);
}

// Which table rows the observation AFTER this step shows: known before the step runs -- next_chunk + 1 while the episode
// goes on (simulator.py:105-106), the first chunk of the next sample when it ends and the env resets itself
// (mansy_env.py:100-101, simulator.py:45), the last chunk again for a terminal observation (mansy_env.py:208-223).
__device__ __forceinline__ void next_obs_chunk(const SimDev &S, const EnvState &st, int auto_reset, int &video, int &pair,
                                               int &start_chunk, int &chunk) {
  video = st.video; pair = st.pair; start_chunk = st.start_chunk;
  if (st.flags & kFlagFinished) { chunk = min(st.next_chunk, st.end_chunk); return; }
  if (st.next_chunk + 1 > st.end_chunk && auto_reset) {
    const int4 q0 = __ldg(reinterpret_cast<const int4 *>(S.ep_init + st.cursor % S.n_samples));     // video, pair, trace, end_chunk
    video = q0.x;
    pair = q0.y;
    start_chunk = __ldg(&S.ep_init[st.cursor % S.n_samples].start_chunk);
    chunk = min(S.startup_download + 1, q0.w);
    return;
  }
  chunk = min(st.next_chunk + 1, st.end_chunk);
}

// What a step reads before it knows the action: the actual-viewport mask, the pyramid scales of this lane's tile
// row, the prediction accuracy, this lane's entry of the bandwidth-trace window and the trace length.  The fused
// rollout kernel issues these loads while the policy is still sampling.
#ifdef MANSY_STEP_PROFILE   // tuning builds only: SM-clock stamps of one thread inside step_env
static __device__ long long *d_step_prof = nullptr;
#define MANSY_STEP_STAMP(k, dep) do { if (d_step_prof && blockIdx.x == 0 && threadIdx.x == 256) d_step_prof[k] = clock64() + ((dep) ? 0 : 0); } while (0)
#else
#define MANSY_STEP_STAMP(k, dep) do { } while (0)
#endif

struct StepInputs {
  uint64_t gt;
  uint32_t scales;
  int tlen;
  double acc, win, rwin;
  const double *tr, *trr;
  size_t vi;
  // outcome-table path of the fused kernel: shared-memory address of this (viewport pair, chunk)'s 16 entries, copied there
  // (cp.async) before the action is known; 0 = read the one entry from the table once it is
  uint32_t outcomes_smem;
};

__device__ __forceinline__ StepInputs step_prefetch(const SimDev &S, const EnvState &st, int sub) {
  StepInputs in;
  const size_t vi = (size_t)st.pair * S.n_vp_chunks + (st.next_chunk - st.start_chunk);   // hmdtrace.py:16-19
  in.vi = vi;
  in.gt = __ldg(S.vp_gt + vi);
  // pyramid scales (toroidal Chebyshev distance to the predicted viewport, utils/common.py:142-168) of this lane's
  // 8 tiles, 4 bits each: a table derived from vp_pred when the handle is created (same tile_scale_masks code)
  in.scales = __ldg(S.vp_scale + vi * 8 + sub);
  in.acc = __ldg(S.vp_acc + vi);
  // bandwidth-trace window and length: issued before the dependent gather / reduction work
  in.tr = S.trace + (size_t)st.trace * S.trace_stride;
  in.trr = S.trace_rcp + (size_t)st.trace * S.trace_stride;
  in.win = __ldg(in.tr + st.cur_idx + sub);
  in.rwin = __ldg(in.trr + st.cur_idx + sub);
  in.tlen = __ldg(S.trace_len + st.trace);
  in.outcomes_smem = 0;
  return in;
}

struct ChunkParts {
  int sz;              // chunk bytes (simulator.py:100)
  double q1, intra;    // qoe.py:23-28
};

// The gather form of (chunk, action) -> parts (utils/common.py:101-193, simulator.py:94-101, qoe.py:23-28), 8 lanes
// per chunk: lane `sub` owns tiles 8*sub .. 8*sub+7.  Steps use it when no outcome table exists (or tile versions are
// asked for) and outcome_build_kernel uses it to fill the table, so both paths produce the same bits.
__device__ __forceinline__ ChunkParts chunk_parts_gather(const SimDev &S, int video, int c, uint64_t gt, uint32_t scales, int rin,
                                                         int rout, int sub, unsigned gmask, uint8_t *__restrict__ ver_row) {
  const uint32_t vtab = (S.lut[rout] & ~7u) | (uint32_t)rin;   // 3-bit entries: scale 0 -> rate_in, s >= 1 -> lut[rate_out][s]
  const uint32_t gbyte = (uint32_t)(gt >> (8 * sub)) & 0xFFu;  // actual-viewport bits of this lane's tiles
  const size_t tab = ((size_t)video * S.n_chunks + c) * kTableRow;
  int sz = 0;
  double mq = 0.0;
  float q[8];
  uint32_t vpack_lo = 0, vpack_hi = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = sub * 8 + i;
    const int ver = (int)((vtab >> (3u * ((scales >> (4 * i)) & 7u))) & 7u);
    const int off = ver * kTiles + t;
    sz += __ldg(S.size + tab + off);
    q[i] = __ldg(S.quality + tab + off);
    if ((gbyte >> i) & 1u) mq = dadd(mq, (double)q[i]);
    if (i < 4) vpack_lo |= (uint32_t)ver << (8 * i); else vpack_hi |= (uint32_t)ver << (8 * (i - 4));
  }
  if (ver_row) reinterpret_cast<uint2 *>(ver_row)[sub] = make_uint2(vpack_lo, vpack_hi);
  ChunkParts cp;
  cp.sz = group_sum(sz, gmask);
  mq = group_sum(mq, gmask);
  const int cnt = __popcll(gt);
  const double sm = (double)cnt, rsm = S.rcp_count[cnt];
  // qoe.py:22-28 (float64 chain; |q - vq| is evaluated in float32 like the reference's array op)
  const double vq = ddiv_rcp(mq, sm, rsm);
  const float vq32 = (float)vq;
  double dev = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if ((gbyte >> i) & 1u) dev = dadd(dev, (double)fabsf(fsub(q[i], vq32)));
  dev = group_sum(dev, gmask);
  cp.intra = ddiv_rcp(ddiv_rcp(dev, sm, rsm), S.max_quality, S.rcp_max_quality);
  cp.q1 = ddiv_rcp(vq, S.max_quality, S.rcp_max_quality);
  return cp;
}

// One chunk-step of one environment (8 lanes).  Returns the reward; `over` tells whether the
// episode ended.  aux_row / ver_row may be NULL.
__device__ __forceinline__ double step_env(const SimDev &S, EnvState &st, float (&slot)[8], int sub, unsigned gmask,
                                           int action, bool &over, double *__restrict__ aux_row,
                                           uint8_t *__restrict__ ver_row, const StepInputs &in) {
  MANSY_STEP_STAMP(0, action);
  const int c = st.next_chunk;
  const uint64_t gt = in.gt;
  const double acc = in.acc;

  int rin, rout;
  action_to_rates(action, rin, rout);
  // (chunk, action) -> chunk bytes, viewport quality, intra-chunk variance: from the outcome table when the handle
  // has one (and the caller does not want the per-tile versions), else gathered here
  ChunkParts cp;
  if (S.outcome != nullptr && ver_row == nullptr) {
    const int a16 = (action >= 0 && action < kActions) ? action : kActions;      // 15 = out-of-table action, rates (0, 0)
    if (in.outcomes_smem) {
      const float4 x = lds128(in.outcomes_smem + (uint32_t)a16 * 32u), y = lds128(in.outcomes_smem + (uint32_t)a16 * 32u + 16u);
      cp.q1 = f2_as_double(x.x, x.y);
      cp.intra = f2_as_double(x.z, x.w);
      cp.sz = (int)__float_as_uint(y.x);
    } else {
      const double2 *o = reinterpret_cast<const double2 *>(S.outcome + in.vi * kOutcomeActions + a16);
      const double2 qa = __ldg(o);
      cp.q1 = qa.x; cp.intra = qa.y;
      cp.sz = __ldg(reinterpret_cast<const int *>(o + 1));
    }
  } else {
    cp = chunk_parts_gather(S, st.video, c, gt, in.scales, rin, rout, sub, gmask, ver_row);
  }
  const int sz = cp.sz;
  MANSY_STEP_STAMP(1, sz);

  // network.py:22-35 / buffer.py:8-15
  bool ok = true;
  const double dl = trace_download_window((double)sz, in.win, in.rwin, in.tr, in.trr, in.tlen, st.cur_idx, st.cur_time, sub, gmask, ok);
  if (!ok && sub == 0) atomicExch(S.error_flag, 1);
  MANSY_STEP_STAMP(2, sz);
  MANSY_STEP_STAMP(3, dl > 0.0);
  const double rebuf = buffer_push(st.buf, S.chunk_length, dl);

  const QoE r = qoe_from_parts(cp.q1, cp.intra, rebuf, st.ep_step == 0, st.prev_vq, (double)st.w0, (double)st.w1, (double)st.w2);
  MANSY_STEP_STAMP(4, r.qoe > 0.0);
  double reward = r.qoe;
  if (S.reward_mode == MANSY_REWARD_QOE_NORM)
    reward = ddiv(r.qoe, dadd(dadd((double)st.w0, (double)st.w1), (double)st.w2));
  st.sum_qoe = dadd(st.sum_qoe, r.qoe);
  st.sum_q1 = dadd(st.sum_q1, r.q1);
  st.sum_q2 = dadd(st.sum_q2, r.q2);
  st.sum_q3 = dadd(st.sum_q3, r.q3);
  st.ep_return = dadd(st.ep_return, reward);

  // mansy_env.py:192-206: push the newest history values (ring slot = episode step & 7)
  if (sub == (st.ep_step & 7)) {
    slot[0] = (float)ddiv_rcp(ddiv((double)sz, dl), S.max_throughput, S.rcp_max_throughput);
    slot[1] = S.rate_norm_hist[rin];
    slot[2] = S.rate_norm_hist[rout];
    slot[3] = (float)acc;
    slot[4] = (float)r.q1;
    slot[5] = (float)r.q3;
    slot[6] = (float)ddiv_rcp(r.q2, S.startup_d, S.rcp_startup_d);
    slot[7] = (float)r.q2;
  }
  MANSY_STEP_STAMP(5, slot[0] > 0.f);
  st.ep_step += 1;
  st.next_chunk = c + 1;                                   // simulator.py:105-106
  over = st.next_chunk > st.end_chunk;
  const int la = (action >= 0 && action < kActions) ? action : kActions;
  st.flags = (st.flags & 0xFF) | (la << 8);

  if (aux_row) {
    double a0, a1;
    switch (sub) {
      case 0: a0 = (double)sz; a1 = dl; break;
      case 1: a0 = rebuf; a1 = st.buf; break;
      case 2: a0 = (double)st.cur_idx; a1 = st.cur_time; break;
      case 3: a0 = r.qoe; a1 = r.q1; break;
      case 4: a0 = r.q2; a1 = r.q3; break;
      case 5: a0 = (double)st.next_chunk; a1 = (double)st.ep_step; break;
      case 6: a0 = (double)st.sample_id; a1 = reward; break;
      default: a0 = (double)(uint32_t)(gt & 0xFFFFFFFFULL); a1 = (double)(uint32_t)(gt >> 32); break;
    }
    reinterpret_cast<double2 *>(aux_row)[sub] = make_double2(a0, a1);
  }
  return reward;
}

// envs/mansy_env.py:271-290: what `_log` records when an episode ends, plus running totals.  Lane `sub` of the
// environment's 8 lanes writes one "last episode" column and updates one "totals" column (same additions as a single
// thread walking the row, spread over the lanes so that an episode end costs one load round trip, not seven).
__device__ __forceinline__ void finish_episode(const SimDev &S, int e, const EnvState &st, int sub) {
  double *row = S.stats + (size_t)e * MANSY_STATS_DOUBLES;
  double last = 0.0, add = 0.0;
  int lc = -1, tc = -1;
  switch (sub) {
    case 0: last = add = st.sum_qoe; lc = MANSY_STAT_LAST_SUM_QOE; tc = MANSY_STAT_TOT_SUM_QOE; break;
    case 1: last = add = st.sum_q1; lc = MANSY_STAT_LAST_SUM_QOE1; tc = MANSY_STAT_TOT_SUM_QOE1; break;
    case 2: last = add = st.sum_q2; lc = MANSY_STAT_LAST_SUM_QOE2; tc = MANSY_STAT_TOT_SUM_QOE2; break;
    case 3: last = add = st.sum_q3; lc = MANSY_STAT_LAST_SUM_QOE3; tc = MANSY_STAT_TOT_SUM_QOE3; break;
    case 4: last = add = (double)st.ep_step; lc = MANSY_STAT_LAST_STEPS; tc = MANSY_STAT_TOT_STEPS; break;
    case 5: last = (double)st.sample_id; add = 1.0; lc = MANSY_STAT_LAST_SAMPLE; tc = MANSY_STAT_TOT_EPISODES; break;
    case 6: last = add = st.ep_return; lc = MANSY_STAT_LAST_RETURN; tc = MANSY_STAT_TOT_RETURN; break;
    default: break;
  }
  if (lc >= 0) {
    row[lc] = last;
    row[tc] += add;
  }
}


}  // namespace mansy

namespace mansy {
__device__ __forceinline__ double step_env(const SimDev &S, EnvState &st, float (&slot)[8], int sub, unsigned gmask,
                                           int action, bool &over, double *__restrict__ aux_row,
                                           uint8_t *__restrict__ ver_row) {
  return step_env(S, st, slot, sub, gmask, action, over, aux_row, ver_row, step_prefetch(S, st, sub));
}
}  // namespace mansy
