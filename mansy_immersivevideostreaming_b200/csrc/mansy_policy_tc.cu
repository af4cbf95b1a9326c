// mansy_policy_tc.cu -- tcgen05 (5th-gen tensor core) policy / value forward for sm_100a.
//
// Reference: bitrate_selection/models/mansy.py:26-51,63-66,77-80, models/simple_rl.py:21-35,46-49,60-63,
// Categorical sampling run_mansy.py:228-229.  The reference runs these layers in TF32
// (torch.set_float32_matmul_precision('high'), run_mansy.py:253); this kernel uses kind::tf32 MMAs
// with fp32 accumulation in tensor memory.
//
// One CTA = one tile of 128 environments (persistent over tiles).  The fp32 observation rows are
// the A operand exactly as the simulator wrote them: TMA brings [128 rows x 32 floats] boxes
// (SWIZZLE_128B) of the row into shared memory, so there is no conversion or gather pass.  Every
// branch of the FeatureNet starts on a multiple of 8 floats (= one TF32 K-step of 32 bytes), so
// layer 1 is, per branch, a run of M128 x N128 x K8 MMAs over the K-steps the branch owns,
// accumulated in a ping-pong TMEM tile D1.  Four epilogue warps add the bias, apply LeakyReLU and
// store the 128x128 feature tile back IN PLACE in tensor memory (tcgen05.st), where it is the A
// operand of layer 2 (A-from-TMEM MMA): D2[128 x 256] += feat_b * Wfc_b^T (actor.fc and critic.fc
// stacked; the shared FeatureNet is evaluated once).  TMEM: 2 x 128 + 256 = 512 columns; shared
// memory is TMA staging (5 x 32 KB) plus the resident head weights.  What bounds the kernel is the
// per-SM TMA ingest (measured ~43 B/clk): a CTA streams 2.4 MB of operands per tile.  The hidden
// activations (bias, LeakyReLU, residual) are stored back over D2 and the heads (128 -> 15,
// 128 -> 1) are one more A-from-TMEM MMA chain, D3[128 x 16] = hid[128 x 256] * Wout^T, with the
// block-structured [16 x 256] head matrix resident in shared memory; softmax and the categorical
// sample run on its 16 columns.
//
// Warp roles: warps 0..5 = TMA producers (two per stage, one per 32 KB half: a thread's TMA issues
// serialise at ~370 cycles each, measured with tools/ubench/tma_ingest.cu, so one producer cannot
// feed the tensor core), warp 6 = TMEM allocator + MMA issuer (one elected lane), warps 8..15 = epilogue
// (TMEM lane quarter = warp & 3, two warps per quarter splitting the columns).  Producers and
// issuer walk the same host-built job list (c_tc[slot].jobs); shared-memory stages are recycled
// through full/empty mbarriers, accumulators through tcgen05.commit barriers.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "mansy_policy.cuh"
#include "mansy_step.cuh"
#include "mansy_tc.cuh"

namespace mansy {

const SimDev *sim_dev_of(mansy_handle_t h);     // mansy_sim.cu

constexpr int kTcSlots = 3;            // policies with live tensor-core state per process
constexpr int kTcMaxJobs = 96;
constexpr int kTcStages = 3;
constexpr int kTcProducers = 2 * kTcStages;   // warps 0..5: producer w owns half (w & 1) of stage w >> 1.  (A stage
                                              // refilled by different producers lets one run two phases ahead,
                                              // where parity waits alias -- hence the fixed ownership.)
constexpr int kTcMmaWarp = 6;                 // warp 7 idles
constexpr int kTcEpiWarp0 = 8;                // warps 8..15
constexpr int kTcThreads = 32 * (kTcEpiWarp0 + 8);
constexpr uint32_t kStageBytes = 65536;   // L1 job: up to 2 A boxes + 2 W1 boxes of 16 KB;  L2 job: 2 Wfc boxes of 32 KB
constexpr uint32_t kBoxBytes = 16384;     // 128 rows x 128 B
constexpr uint32_t kWoutBytes = 16384;     // head matrix [16 rows x 256 K] fp32 = 8 boxes of [16 x 32 floats]
constexpr uint32_t kD3RecvBytes = 8192;    // cluster kernel: head partials of this CTA's 32 rows from the 4 ranks
// cluster kernel: the WHOLE head matrix (8 boxes [16 x 32 floats]) + the residual head matrix (4 boxes) + the sampled actions
constexpr uint32_t kClWoutBytes = 16384 + 8192 + 256;
constexpr uint32_t kTcSmemBytesSolo = kTcStages * kStageBytes + kWoutBytes + 256 /*barriers*/ + kD3RecvBytes + 1024 /*alignment*/;
constexpr uint32_t kTcSmemBytesCluster = kTcStages * kStageBytes + kClWoutBytes + 256 /*barriers*/ + 1024 /*alignment*/;
constexpr uint32_t kTcSmemBytes = kTcSmemBytesSolo > kTcSmemBytesCluster ? kTcSmemBytesSolo : kTcSmemBytesCluster;

enum : uint8_t { kJobL1 = 0, kJobL2 = 1 };
enum : uint8_t { kFlagFirst = 1, kFlagLast = 2, kFlagTileFirstL2 = 4, kFlagTileLastL2 = 8 };

struct TcJob {
  uint8_t type;     // kJobL1 / kJobL2
  uint8_t a_box;    // L1: first float of the first observation box (32 floats wide), in units of 4 floats   L2: -
  uint8_t w_box;    // L1: first column of the first box of the layer-1 weight image, same units               L2: -
  uint8_t s_lo;     // L1: first K-step inside the job (0..7)      L2: -
  uint8_t s_hi;     // L1: one past the last K-step (<= 8)         L2: -
  uint8_t slot;     // branch in processing order (D1 / feature buffer = slot & 1)
  uint8_t flags;
  uint8_t chunk;    // L1: number of boxes (1 or 2)   L2: first of two 32-float chunks of the branch's 128 features (0, 2)
};

// Split-K plan of the 4-CTA cluster kernel: CTA `rank` of a cluster owns a subset of the branches (layer 1 of
// the branch and the rows of actor.fc / critic.fc that multiply its 128 features).
constexpr int kTcRanks = 4;
constexpr int kTcRankJobs = 32;
struct TcRankPlan {
  TcJob jobs[kTcRankJobs];     // job.slot = LOCAL branch index (D1 buffer = slot & 1)
  int32_t n_jobs, n_branches;
  int32_t resid_local;         // local index of the residual branch (always the last one) or -1
  int32_t pad;
  uint8_t branch[16];          // local index -> processing-order index (bias1 row, Wfc column block)
};

struct TcConst {
  float bias2[2 * kHidden];    // actor.fc bias | critic.fc bias
  float bias1[kMaxBranches][kHidden];   // processing order
  float bout[16];
  TcJob jobs[kTcMaxJobs];
  int32_t n_jobs, n_branches, residual_slot, softmax;
  TcRankPlan rank[kTcRanks];
  // "memo" plans: the 320-input table branches (next chunk sizes / qualities) are left out -- their contribution to
  // the hidden pre-activation comes from the (video, chunk) table of policy_memo_for.  rank_memo: the split-K cluster
  // kernel; solo_memo: the one-CTA-per-tile kernel (job.slot / branch[] as in TcRankPlan).
  TcRankPlan rank_memo[kTcRanks];
  TcRankPlan solo_memo;
};

__constant__ TcConst c_tc[kTcSlots];

struct TcArgs {
  int32_t n, n_tiles;
  float *logits;       // [n][16] or NULL
  float *value;        // [n] or NULL
  int32_t *actions;    // [n] or NULL
  float *logp;         // [n] or NULL
  float *feat_dbg;     // [n][n_branches*128] (processing order) or NULL
  float *hid_dbg;      // [n][256] or NULL
  uint64_t seed;
  int64_t step;
  int32_t env_offset;
  long long *timeline;  // debug: clock64 stamps [4][128] (producer issue, data arrival, mma committed, epilogue) of one CTA or NULL
  int32_t timeline_cta; // blockIdx.x of the CTA that writes the timeline
  float4 *scratch;      // cluster kernel: [tile][dst rank 4][src rank 4][column half 2][float4 column 8][row 128] partial exchange (L2)
  const float *memo;    // [video * n_chunks + chunk][256] or NULL (all branches through the tensor pipe)
  const EnvState *memo_state;   // row i of the observations is the current observation of environment i of this array
  int32_t memo_n_chunks;
};

// Row of the memo table for environment `e`: the chunk its current observation describes (emit_obs: min(next_chunk,
// end_chunk), mansy_env.py:208-230) of its episode's video.
__device__ __forceinline__ int memo_row_of(const EnvState *st, int e, int n_chunks) {
  // (L2 loads: inside the fused rollout kernel the records change from step to step)
  const int video = __ldcg(&st[e].video), nc = __ldcg(&st[e].next_chunk), ec = __ldcg(&st[e].end_chunk);
  return video * n_chunks + min(nc, ec);
}

struct TcState {
  int slot = -1;
  CUtensorMap map_w1, map_wfc, map_wout, map_wres;
  int obs_floats = 0;          // 784 / 400
  int n_branches = 0;
  int split = 0;               // 0 = by batch size, 1 = one CTA per 128-env tile, 4 = split-K cluster of 4 CTAs per tile
  float4 *scratch = nullptr;   // partial-sum exchange of the cluster kernel (128 KB per rank and tile), grown on demand
  int scratch_tiles = 0;
  // observation tensor maps are encoded once per (pointer, rows, stride): a rollout buffer is reused call after call
  struct ObsMap { const float *ptr = nullptr; uint64_t rows = 0, stride = 0; CUtensorMap map; } obs_maps[4];
  int obs_map_next = 0;
};


// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
template <int SLOT>
__global__ void __launch_bounds__(kTcThreads, 1)
policy_tc_kernel(const __grid_constant__ CUtensorMap map_obs, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_wfc, const __grid_constant__ CUtensorMap map_wout,
                 const TcArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const TcConst &K = c_tc[SLOT];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage0 = base;
  const uint32_t wout_s = base + kTcStages * kStageBytes;
  const uint32_t bars = wout_s + kWoutBytes;
  // barrier map (8 bytes each)
  // A stage has one "full" barrier per 32 KB half: 65536 pending transaction bytes on one mbarrier fault on
  // sm_100a (compute-sanitizer flags the arrive.expect_tx), and the issuer can start on the first half earlier.
  const uint32_t bar_full = bars;                    // [kTcStages][2] TMA -> MMA
  const uint32_t bar_empty = bars + 64;              // [kTcStages]    MMA (commit) -> TMA
  const uint32_t bar_d1_full = bars + 128;           // [2]           MMA (commit) -> epilogue
  const uint32_t bar_feat_full = bars + 144;         // [2]           epilogue (128 arrivals) -> MMA
  const uint32_t bar_d2_full = bars + 160;           //               MMA (commit) -> epilogue
  const uint32_t bar_d2_empty = bars + 168;          //               epilogue (128 arrivals) -> MMA
  const uint32_t bar_hid_full = bars + 176;          //               epilogue (128 arrivals) -> MMA: hid stored over D2
  const uint32_t bar_d3_full = bars + 184;           //               MMA (commit) -> epilogue: head outputs ready
  const uint32_t bar_wout = bars + 192;              //               TMA -> MMA: head matrix resident (once)
  const uint32_t tmem_slot = bars + 200;             // uint32 written by tcgen05.alloc

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(bar_full + 16 * s, 1);
      mbar_init(bar_full + 16 * s + 8, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_d1_full + 8 * b, 1);
      mbar_init(bar_feat_full + 8 * b, 256);
    }
    mbar_init(bar_d2_full, 1);
    mbar_init(bar_d2_empty, 128);
    mbar_init(bar_hid_full, 256);
    mbar_init(bar_d3_full, 1);
    mbar_init(bar_wout, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_obs) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wfc) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wout) : "memory");
  }
  if (warp == kTcMmaWarp) {   // TMEM: all 512 columns (D1[0] 0..127, D1[1] 128..255, D2 256..511)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const bool memo = A.memo != nullptr;
  const TcJob *const jobs = memo ? K.solo_memo.jobs : K.jobs;
  const int n_jobs = memo ? K.solo_memo.n_jobs : K.n_jobs;
  if (A.timeline && blockIdx.x == 0 && threadIdx.x == 0) A.timeline[511] = clock64();
  griddep_launch();

  if (warp < kTcProducers) {
    // ===== TMA producers: job `it` uses stage it % kTcStages; producers 2s and 2s+1 fill its two halves =====
    const int half = warp & 1;
    {
      if (warp == 0 && elect_one()) {
        mbar_expect_tx(bar_wout, kWoutBytes);          // head matrix: resident for the whole kernel
        for (int b = 0; b < 8; ++b) tma_load_2d(wout_s + b * 2048, &map_wout, b * 32, 0, bar_wout);
      }
      __syncwarp();
      griddep_wait();                                  // observation rows come from the previous kernel in the stream
      uint32_t it = 0, s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        for (int j = 0; j < n_jobs; ++j, ++it, s = (s + 1 == kTcStages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          if ((int)s != (warp >> 1)) continue;
          const TcJob job = jobs[j];
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t dst = stage0 + s * kStageBytes, full = bar_full + 16 * s + 8 * half;
          if (elect_one()) {
            if (job.type == kJobL1) {     // stage: [A box 0][A box 1][W1 box 0][W1 box 1]
              if (half < job.chunk) {
                mbar_expect_tx(full, 2 * kBoxBytes);
                tma_load_2d(dst + half * kBoxBytes, &map_obs, job.a_box * 4 + half * 32, tile * 128, full);
                tma_load_2d(dst + (2 + half) * kBoxBytes, &map_w1, job.w_box * 4 + half * 32, 0, full);
              } else {
                mbar_arrive(full);        // nothing to load: keep both halves' phases in step
              }
            } else {                      // stage: [Wfc chunk c][Wfc chunk c + 1], 256 rows x 128 B each
              if (half < job.s_hi) {
                mbar_expect_tx(full, 2 * kBoxBytes);
                tma_load_2d(dst + half * 2 * kBoxBytes, &map_wfc,
                            (memo ? (int)K.solo_memo.branch[job.slot] : (int)job.slot) * kHidden + (job.chunk + half) * 32, 0, full);
              } else {
                mbar_arrive(full);
              }
            }
          }
          __syncwarp();
          if (A.timeline && blockIdx.x == 0 && it < 128 && lane == 0 && half == 0) A.timeline[it] = clock64();
        }
      }
    }
  } else if (warp == kTcMmaWarp) {
    // ===== MMA issuer =====
    {
      constexpr uint32_t kIdesc128 = idesc_tf32(128), kIdesc256 = idesc_tf32(256), kIdesc16 = idesc_tf32(16);
      uint32_t it = 0, feat_use[2] = {0, 0}, tile_i = 0, s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_i) {
        // the head epilogue of the previous tile reads D2 (hid), D1[1] (residual) and D3 (= D1[0] columns 0..15)
        if (tile_i > 0) mbar_wait(bar_d2_empty, (tile_i - 1) & 1u);
        TcJob job = jobs[0];
        for (int j = 0; j < n_jobs; ++j, ++it, s = (s + 1 == kTcStages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          const TcJob next_job = jobs[j + 1 < n_jobs ? j + 1 : 0];      // fetched before the waits below
          const uint32_t buf = job.slot & 1u;
          const uint32_t st_addr = stage0 + s * kStageBytes;
          if (job.type == kJobL1) {
            // D1[buf] still holds the features of branch slot-2 until that branch's layer-2 MMAs have
            // read them; those precede this job in the issue order and the tensor pipe runs in order.
            const uint32_t d1 = tmem_base + buf * 128u;
            const uint32_t a_lo = smem_desc_lo(st_addr), b_lo = smem_desc_lo(st_addr + 2 * kBoxBytes);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mbar_wait(bar_full + 16 * s + 8 * h, ph);
              tc_fence_after();
              if (h == 0 && A.timeline && blockIdx.x == 0 && lane == 0) A.timeline[128 + j] = clock64();
              if (elect_one()) {
#pragma unroll
                for (int st = 4 * h; st < 4 * h + 4; ++st)
                  if (st >= job.s_lo && st < job.s_hi)
                    umma_tf32(d1, make_desc(a_lo + h * (kBoxBytes >> 4) + 2 * (st & 3)),
                              make_desc(b_lo + h * (kBoxBytes >> 4) + 2 * (st & 3)), kIdesc128,
                              ((job.flags & kFlagFirst) && st == job.s_lo) ? 0u : 1u);
              }
              __syncwarp();
            }
            if (elect_one()) {
              umma_commit(bar_empty + 8 * s);
              if (job.flags & kFlagLast) umma_commit(bar_d1_full + 8 * buf);
            }
            __syncwarp();
            if (A.timeline && blockIdx.x == 0 && lane == 0) A.timeline[256 + j] = clock64();
          } else {
            if (job.chunk == 0) {   // feature tile of this branch written by the epilogue warps
              mbar_wait(bar_feat_full + 8 * buf, feat_use[buf] & 1u);
              ++feat_use[buf];
            }
            const uint32_t d2 = tmem_base + 256u;
            const uint32_t fa = tmem_base + buf * 128u + job.chunk * 32u;     // features live in D1[buf]
            const uint32_t b_lo = smem_desc_lo(st_addr);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mbar_wait(bar_full + 16 * s + 8 * h, ph);
              tc_fence_after();
              if (h == 0 && A.timeline && blockIdx.x == 0 && lane == 0) A.timeline[128 + j] = clock64();
              if (elect_one()) {
#pragma unroll
                for (int ks = 4 * h; ks < 4 * h + 4; ++ks)
                  if (h < job.s_hi)
                    umma_tf32_ts(d2, fa + ks * 8, make_desc(b_lo + h * (2 * kBoxBytes >> 4) + 2 * (ks & 3)), kIdesc256,
                                 ((job.flags & kFlagTileFirstL2) && ks == 0) ? 0u : 1u);
              }
              __syncwarp();
            }
            if (elect_one()) {
              umma_commit(bar_empty + 8 * s);
              if (job.flags & kFlagTileLastL2) umma_commit(bar_d2_full);
            }
            __syncwarp();
            if (A.timeline && blockIdx.x == 0 && lane == 0) A.timeline[256 + j] = clock64();
          }
          job = next_job;
        }
        // heads: D3[128 x 16] = hid[128 x 256] (stored over D2 by the epilogue warps) * Wout^T
        if (tile_i == 0) mbar_wait(bar_wout, 0);
        mbar_wait(bar_hid_full, tile_i & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = smem_desc_lo(wout_s);
#pragma unroll
          for (int ks = 0; ks < 32; ++ks)
            umma_tf32_ts(tmem_base, tmem_base + 256u + ks * 8, make_desc(w_lo + (ks >> 2) * (2048 >> 4) + (ks & 3) * 2),
                         kIdesc16, ks > 0 ? 1u : 0u);
          umma_commit(bar_d3_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kTcEpiWarp0) {
    // ===== epilogue warps (8..15): two warps per TMEM lane quarter, `half` selects the columns =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - kTcEpiWarp0) >> 2;
    const int r = q * 32 + lane;            // row (environment) inside the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int nb = memo ? K.solo_memo.n_branches : K.n_branches;
    uint32_t d1_use[2] = {0, 0}, tile_i = 0;
    griddep_wait();
    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_i) {
      const int env = tile * 128 + r;
      const bool live = env < A.n;
      const float4 *memo_row = memo ? reinterpret_cast<const float4 *>(A.memo + (size_t)memo_row_of(A.memo_state, live ? env : 0, A.memo_n_chunks) * 256)
                                    : nullptr;
      for (int i = 0; i < nb; ++i) {
        const uint32_t buf = i & 1u;
        const int bi = memo ? (int)K.solo_memo.branch[i] : i;      // processing-order index: bias row, feat_dbg block
        mbar_wait(bar_d1_full + 8 * buf, d1_use[buf] & 1u);
        ++d1_use[buf];
        tc_fence_after();
        if (A.timeline && blockIdx.x == 0 && threadIdx.x == 32 * kTcEpiWarp0 && tile_i == 0) A.timeline[384 + 2 * i] = clock64();
#pragma unroll 1
        for (int c = 2 * half; c < 2 * half + 2; ++c) {
          float v[32];
          const uint32_t ta = tmem_base + lane_addr + buf * 128u + c * 32u;
          tmem_ld32(ta, v);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = leaky(v[jj] + K.bias1[bi][c * 32 + jj]);
          tmem_st32(ta, v);        // features replace the accumulator in place: A operand of layer 2
          if (A.feat_dbg && live) {
            float *dst = A.feat_dbg + (size_t)env * (K.n_branches * kHidden) + bi * kHidden + c * 32;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) dst[jj] = v[jj];
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_feat_full + 8 * buf);
        if (A.timeline && blockIdx.x == 0 && threadIdx.x == 32 * kTcEpiWarp0 && tile_i == 0) A.timeline[384 + 2 * i + 1] = clock64();
      }

      // ---- final epilogue: hidden layer activation, residual, heads, sample -----------------
      mbar_wait(bar_d2_full, tile_i & 1u);
      tc_fence_after();
      if (A.timeline && blockIdx.x == 0 && threadIdx.x == 32 * kTcEpiWarp0 && tile_i == 0) A.timeline[384 + 2 * nb] = clock64();
      const int rs = memo ? K.solo_memo.resid_local : K.residual_slot;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float va[32], rr[32];
        float4 mm[8];
        if (memo) {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) mm[c4] = __ldg(memo_row + half * 32 + c * 8 + c4);
        }
        if (rs >= 0) {           // residual: the features of the last branch are still in D1[rs & 1]
          tmem_ld32(tmem_base + lane_addr + (uint32_t)(rs & 1) * 128u + c * 32u, rr);
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) rr[jj] = 0.f;
        }
        {                                          // half 0: actor.fc columns, half 1: critic.fc columns
          const uint32_t ta = tmem_base + lane_addr + 256u + half * 128u + c * 32u;
          tmem_ld32(ta, va);
          if (memo) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              va[4 * c4] += mm[c4].x; va[4 * c4 + 1] += mm[c4].y; va[4 * c4 + 2] += mm[c4].z; va[4 * c4 + 3] += mm[c4].w;
            }
          }
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)          // mansy.py:65,79: fc(features) + qoe_features
            va[jj] = leaky(va[jj] + K.bias2[half * kHidden + c * 32 + jj]) + rr[jj];
          tmem_st32(ta, va);
          if (A.hid_dbg && live) {
            float *dst = A.hid_dbg + (size_t)env * 256 + half * kHidden + c * 32;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) dst[jj] = va[jj];
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_hid_full);
      if (half != 0) continue;        // the 16 head outputs of a row are handled by one thread
      mbar_wait(bar_d3_full, tile_i & 1u);
      tc_fence_after();
      float acc[16];
      tmem_ld16(tmem_base + lane_addr, acc);
      tc_fence_before();
      mbar_arrive(bar_d2_empty);      // D2, the residual and D3 are consumed: the next tile may start
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[o] += K.bout[o];
      if (live) {
        float p[kActions];
#pragma unroll
        for (int o = 0; o < kActions; ++o) p[o] = acc[o];
        if (K.softmax == 2) {     // QoE identifier: torch.sigmoid on its 3 outputs (mansy.py:141)
#pragma unroll
          for (int o = 0; o < 3; ++o) p[o] = 1.0f / (1.0f + expf(-p[o]));
        } else if (K.softmax) {   // simple_rl.py:48: the actor returns probabilities
          float m = p[0];
#pragma unroll
          for (int o = 1; o < kActions; ++o) m = fmaxf(m, p[o]);
          float s = 0.f;
#pragma unroll
          for (int o = 0; o < kActions; ++o) { p[o] = expf(p[o] - m); s += p[o]; }
#pragma unroll
          for (int o = 0; o < kActions; ++o) p[o] = p[o] / s;
        }
        if (A.value) A.value[env] = acc[15];
        if (A.logits) {
          float4 *dst = reinterpret_cast<float4 *>(A.logits + (size_t)env * 16);
          dst[0] = make_float4(p[0], p[1], p[2], p[3]);
          dst[1] = make_float4(p[4], p[5], p[6], p[7]);
          dst[2] = make_float4(p[8], p[9], p[10], p[11]);
          dst[3] = make_float4(p[12], p[13], p[14], 0.f);
        }
        if (A.actions) {
          int act;
          float lp;
          categorical_sample(p, K.softmax == 1, A.seed, (uint64_t)(A.env_offset + env), (uint64_t)A.step, act, lp);
          A.actions[env] = act;
          if (A.logp) A.logp[env] = lp;
        }
      }
      if (A.timeline && blockIdx.x == 0 && threadIdx.x == 32 * kTcEpiWarp0 && tile_i == 0) A.timeline[384 + 2 * nb + 1] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// split-K cluster kernel (small batches)
// ------------------------------------------------------------------------------------------
// With 4 096 environments there are only 32 tiles of 128 rows, and what bounds a tile is the operand
// stream into ONE SM (2.4 MB of weights + rows at ~50-60 B/clk) -- 32 of 148 SMs busy for ~30 us.  Here a
// cluster of 4 CTAs shares a tile: CTA `rank` runs layer 1 for its own branches and multiplies their
// features with the matching rows of actor.fc | critic.fc, so each CTA streams a quarter of the operands and
// holds a PARTIAL D2[128 x 256] in tensor memory.  The partials are reduce-scattered over distributed shared
// memory (rank d owns hidden columns 64d .. 64d+63: st.shared::cluster into d's idle TMA stages), the owner
// adds bias + LeakyReLU, stores its 64 hidden columns back to TMEM and runs ITS K-slice of the heads
// (D3[128 x 16] partial); the residual `+ qoe_features` (mansy.py:65,79) is linear after the activation, so
// the rank that owns the qoe branch adds feat_qoe * [actor.out ; critic.out]^T to its D3 partial instead of
// shipping the feature tile around.  The four D3 partials are summed on rank 0, which applies the head bias,
// softmax and the categorical sample.  Three barrier.cluster phases: stages idle -> partials delivered ->
// head partials delivered.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
// Exchange buffer of the partial D2 tiles in global memory (it stays in L2): distributed shared memory moves
// only ~20 B/clk per SM (measured: 96 KB of st.shared::cluster took 4 600 - 7 800 cycles), the L2 path several
// times that.  Index in float4: ((((tile * 4 + dst) * 4 + src) * 2 + half) * 8 + c4) * 128 + row.
// (Row split: rank d finishes rows 32d .. 32d+31 of the tile for all 256 hidden columns.)  Index in float4:
// ((tile * 4 + dst) * 4 + src) * 2048 + c4 * 32 + row  with c4 = float4 column 0..63, row = 0..31 of the destination's rows.
constexpr size_t kScratchBlockF4 = 64 * 32;
constexpr size_t kScratchPartialF4 = 4 * 4 * kScratchBlockF4;
// + the memoised table branches of the destination's rows in the same layout: [dst rank 4][c4 64][row 32]
constexpr size_t kScratchMemoF4 = 4 * kScratchBlockF4;
// + the residual's head contribution (feat_qoe * [actor.out ; critic.out]^T, 16 columns) for the destination's rows: [dst 4][c4 4][row 32]
constexpr size_t kScratchResF4 = 4 * 4 * 32;
constexpr size_t kScratchF4PerTile = kScratchPartialF4 + kScratchMemoF4 + kScratchResF4;

// Fused rollout (MODE != MANSY_OBS_NONE): the same cluster also runs the simulator step of its 128 environments
// (32 per CTA, 8 lanes each, on the eight epilogue warps) with the action it has just sampled, writes the next
// observation slab and loops over the rollout steps without leaving the kernel: no launch, drain or refill
// between policy and simulator, and the cluster only synchronises with itself (environments never interact,
// mansy_env.py holds no cross-env state).
struct FusedArgs {
  SimDev S;
  float *obs;          // [slabs][n][obs_stride]   (the TMA map `map_obs` views it as [slabs * n] rows)
  int64_t obs_stride;
  int32_t slabs, n_steps;
  int32_t outcome_prefetch;   // 1: fetch the outcomes of all 16 actions before the action is known; 0: one entry afterwards
  int64_t t0;
  int32_t *actions;    // [slabs][n]
  float *logp, *value, *reward;
  uint8_t *done;
};

#ifdef MANSY_MBAR_WATCHDOG     // debugging builds: every warp leaves (step << 16 | code) in a host-mapped buffer as it moves along
static __device__ volatile int *d_dbg_progress = nullptr;
#define MANSY_DBG(code) do { if (d_dbg_progress && (threadIdx.x & 31) == 0) d_dbg_progress[blockIdx.x * 16 + (threadIdx.x >> 5)] = ((int)k << 16) | (code); } while (0)
#else
#define MANSY_DBG(code) do { } while (0)
#endif

template <int SLOT, int MODE>
__global__ void __cluster_dims__(kTcRanks, 1, 1) __launch_bounds__(kTcThreads, 1)
policy_tc4_kernel(const __grid_constant__ CUtensorMap map_obs, const __grid_constant__ CUtensorMap map_w1,
                  const __grid_constant__ CUtensorMap map_wfc, const __grid_constant__ CUtensorMap map_wout,
                  const __grid_constant__ CUtensorMap map_wres, const TcArgs A, const __grid_constant__ FusedArgs F) {
  constexpr bool kFused = MODE != MANSY_OBS_NONE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const TcConst &K = c_tc[SLOT];
  const uint32_t rank = cluster_ctarank();
  const bool memo = A.memo != nullptr;
  const TcRankPlan &P = memo ? K.rank_memo[rank] : K.rank[rank];
  // A cluster walks the tiles cluster, cluster + n_clusters, ... (one tile per cluster when every tile's cluster is
  // resident -- up to 33 tiles = 4 224 environments on B200 -- several per step beyond that, e.g. two at 8 192)
  const int tile0 = blockIdx.x / kTcRanks, tile_stride = gridDim.x / kTcRanks;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage0 = base;
  const uint32_t wout_s = base + kTcStages * kStageBytes;   // the head matrix [16 x 256]: 8 boxes [16 x 32 floats]
  const uint32_t wres_s = wout_s + 16384;                   // [actor.out ; critic.out] for the residual: 4 boxes
  const uint32_t act_s = wout_s + 24576;                    // fused: the action sampled for each of this CTA's 32 environments
  const uint32_t hid_s = stage0;                            // phase C: the finished hidden rows as the A operand of the heads -- 8 boxes
                                                            // [128 rows x 32 floats] (stages 0 and 1), rows 96..127 of each
  const uint32_t own_s = stage0 + 2 * kStageBytes + 32768;  // phase B -> C: this rank's own partial of its rows, [c4 64][row 32] float4
  const uint32_t sim_s = stage0 + 2 * kStageBytes;          // fused: the third TMA stage, idle between phase A and the next one, stages
                                                            // the simulator phase's loads: [32 records x 128 B][32 x 8 slots x 32 B][32 x 8 lanes x 128 B]
  const uint32_t bars = wout_s + kClWoutBytes;
  const uint32_t bar_full = bars;                    // [kTcStages][2] TMA -> MMA
  const uint32_t bar_empty = bars + 64;              // [kTcStages]    MMA (commit) -> TMA
  const uint32_t bar_d1_full = bars + 128;           // [2]           MMA (commit) -> epilogue
  const uint32_t bar_feat_full = bars + 144;         // [2]           epilogue (256 arrivals) -> MMA
  const uint32_t bar_d2_full = bars + 160;           //               MMA (commit) -> everyone: partial D2 complete
  const uint32_t bar_hid_full = bars + 176;          //               384 arrivals -> MMA: hidden rows stored (shared memory)
  const uint32_t bar_d3_full = bars + 184;           //               MMA (commit) -> spare warp / table-row warp: heads done
  const uint32_t bar_res_full = bars + 240;          //               MMA (commit) -> residual head contribution in tensor memory
  const uint32_t bar_wout = bars + 192;              //               TMA -> MMA: head matrices resident
  const uint32_t tmem_slot = bars + 200;
  const uint32_t bar_tab = bars + 208;               // [4]           bulk table-row loads of 8 environments each (fused)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tl = A.timeline != nullptr && (int)blockIdx.x == (A.timeline_cta & 0xFFFF);     // (stamps of the cluster's last tile of the step win)
  // the rollout step the timeline records: bits 16.. of timeline_cta, default (0) = the second step of a fused launch, a
  // steady-state one; 1 + s = step s (MANSY_TC_TIMELINE_STEP, tools/fused_timeline.py: step 0 is the cold one)
  const int tlk = (A.timeline_cta >> 16) ? (A.timeline_cta >> 16) - 1 : (kFused ? 1 : 0);
  const int n_jobs = P.n_jobs, nloc = P.n_branches, resid = P.resid_local;
  // D3 (head partial, 16 columns) lives in the D1 buffer the residual features do NOT occupy
  const uint32_t d3_col = resid >= 0 ? (uint32_t)((resid & 1) ^ 1) * 128u : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(bar_full + 16 * s, 1);
      mbar_init(bar_full + 16 * s + 8, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_d1_full + 8 * b, 1);
      mbar_init(bar_feat_full + 8 * b, 256);
    }
    mbar_init(bar_d2_full, 1);
    mbar_init(bar_hid_full, 384);
    mbar_init(bar_d3_full, 1);
    mbar_init(bar_res_full, 1);
    mbar_init(bar_wout, 1);
    for (int g = 0; g < 4; ++g) mbar_init(bar_tab + 8 * g, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_obs) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wfc) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wout) : "memory");
  }
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (tl && threadIdx.x == 0) A.timeline[511] = clock64();
  griddep_launch();                // the kernel after us may be scheduled (it waits for our completion itself)

  // epilogue-warp coordinates (used in several phases below)
  const int q = warp & 3;
  const int half = (warp - kTcEpiWarp0) >> 2;
  const int r = q * 32 + lane;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;

  if (warp == 0 && elect_one()) {            // weights do not depend on the previous kernel in the stream
    mbar_expect_tx(bar_wout, 16384u + (resid >= 0 ? 8192u : 0u));
    for (int b = 0; b < 8; ++b) tma_load_2d(wout_s + b * 2048, &map_wout, b * 32, 0, bar_wout);
    if (resid >= 0)
      for (int b = 0; b < 4; ++b) tma_load_2d(wres_s + b * 2048, &map_wres, b * 32, 0, bar_wout);
  }
  __syncwarp();
  if (warp != kTcMmaWarp) griddep_wait();   // rows / state records come from, results go to stream-ordered memory

  // pipeline state carried across the rollout steps of the fused kernel
  uint32_t it = 0, s = 0, ph = 0;          // TMA producers and MMA issuer walk the same job sequence
  uint32_t feat_use[2] = {0, 0}, d1_use[2] = {0, 0};
  const int n_steps = kFused ? F.n_steps : 1;
  uint32_t ts = 0;                          // tile-steps done by this cluster (parity of the once-per-tile-step barriers)
  uint32_t tab_use = 0;                     // phases this lane's bar_tab barrier has been through (a tail tile may skip one)

#pragma unroll 1
  for (int k = 0; k < n_steps; ++k) {
   const int64_t t = kFused ? F.t0 + k : A.step;
   const int cur = kFused ? (int)(t % F.slabs) : 0;
#pragma unroll 1
   for (int tile = tile0; tile < A.n_tiles; tile += tile_stride, ++ts) {
    const uint32_t par = ts & 1u;
    const int env = tile * 128 + r;
    const bool live = env < A.n;
    const int row0 = cur * A.n + tile * 128;          // first row of the tile in the (slab-stacked) observation tensor
    if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[489] = clock64();
    // fused: what the simulator phase of this CTA's 32 environments needs (8 lanes per environment on the epilogue warps)
    EnvState sim_st;
    float sim_slot[8];
    StepInputs sim_in;
    const int sim_et = (int)threadIdx.x - 32 * kTcEpiWarp0;
    const int sim_i = tile * 128 + (int)rank * 32 + (sim_et >> 3);
    const bool sim_live = kFused && warp >= kTcEpiWarp0 && sim_i < A.n;
    uint64_t sim_pred = 0;                       // predicted viewport of the next observation (when the episode goes on)
    // ================= phase A: the rank's branches -> partial D2 =================
    if (warp < kTcProducers) {
      const int hf = warp & 1;
      for (int j = 0; j < n_jobs; ++j, ++it, s = (s + 1 == kTcStages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
        if ((int)s != (warp >> 1)) continue;
        const TcJob job = P.jobs[j];
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t dst = stage0 + s * kStageBytes, full = bar_full + 16 * s + 8 * hf;
        if (elect_one()) {
          if (job.type == kJobL1) {
            if (hf < job.chunk) {
              mbar_expect_tx(full, 2 * kBoxBytes);
              tma_load_2d(dst + hf * kBoxBytes, &map_obs, job.a_box * 4 + hf * 32, row0, full);
              tma_load_2d(dst + (2 + hf) * kBoxBytes, &map_w1, job.w_box * 4 + hf * 32, 0, full);
            } else {
              mbar_arrive(full);
            }
          } else {
            if (hf < job.s_hi) {
              mbar_expect_tx(full, 2 * kBoxBytes);
              tma_load_2d(dst + hf * 2 * kBoxBytes, &map_wfc, (int)P.branch[job.slot] * kHidden + (job.chunk + hf) * 32, 0, full);
            } else {
              mbar_arrive(full);
            }
          }
        }
        __syncwarp();
        if (tl && k == tlk && lane == 0 && hf == 0) A.timeline[j] = clock64();
      }
    } else if (warp == kTcMmaWarp) {
      constexpr uint32_t kIdesc128 = idesc_tf32(128), kIdesc256 = idesc_tf32(256), kIdesc16 = idesc_tf32(16);
      tc_fence_after();
      TcJob job = P.jobs[0];
      for (int j = 0; j < n_jobs; ++j, ++it, s = (s + 1 == kTcStages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
        const TcJob next_job = P.jobs[j + 1 < n_jobs ? j + 1 : 0];
        const uint32_t buf = job.slot & 1u;
        const uint32_t st_addr = stage0 + s * kStageBytes;
        if (job.type == kJobL1) {
          const uint32_t d1 = tmem_base + buf * 128u;
          const uint32_t a_lo = smem_desc_lo(st_addr), b_lo = smem_desc_lo(st_addr + 2 * kBoxBytes);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar_full + 16 * s + 8 * h, ph);
            tc_fence_after();
            if (h == 0 && tl && k == tlk && lane == 0) A.timeline[128 + j] = clock64();
            if (elect_one()) {
#pragma unroll
              for (int st = 4 * h; st < 4 * h + 4; ++st)
                if (st >= job.s_lo && st < job.s_hi)
                  umma_tf32(d1, make_desc(a_lo + h * (kBoxBytes >> 4) + 2 * (st & 3)),
                            make_desc(b_lo + h * (kBoxBytes >> 4) + 2 * (st & 3)), kIdesc128,
                            ((job.flags & kFlagFirst) && st == job.s_lo) ? 0u : 1u);
            }
            __syncwarp();
          }
          if (elect_one()) {
            umma_commit(bar_empty + 8 * s);
            if (job.flags & kFlagLast) umma_commit(bar_d1_full + 8 * buf);
          }
          __syncwarp();
        } else {
          if (job.chunk == 0) {
            mbar_wait(bar_feat_full + 8 * buf, feat_use[buf] & 1u);
            ++feat_use[buf];
          }
          const uint32_t d2 = tmem_base + 256u;
          const uint32_t fa = tmem_base + buf * 128u + job.chunk * 32u;
          const uint32_t b_lo = smem_desc_lo(st_addr);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar_full + 16 * s + 8 * h, ph);
            tc_fence_after();
            if (h == 0 && tl && k == tlk && lane == 0) A.timeline[128 + j] = clock64();
            if (elect_one()) {
#pragma unroll
              for (int ks = 4 * h; ks < 4 * h + 4; ++ks)
                if (h < job.s_hi)
                  umma_tf32_ts(d2, fa + ks * 8, make_desc(b_lo + h * (2 * kBoxBytes >> 4) + 2 * (ks & 3)), kIdesc256,
                               ((job.flags & kFlagTileFirstL2) && ks == 0) ? 0u : 1u);
            }
            __syncwarp();
          }
          if (elect_one()) {
            umma_commit(bar_empty + 8 * s);
            if (job.flags & kFlagTileLastL2) umma_commit(bar_d2_full);
          }
          __syncwarp();
        }
        if (tl && k == tlk && lane == 0) A.timeline[256 + j] = clock64();
        job = next_job;
      }
      if (resid >= 0) {
        // residual through the heads: feat_qoe[128 x 128] * [actor.out ; critic.out]^T for all 128 rows.  The qoe features sit
        // in D1[resid & 1] (their layer-2 MMAs above waited for them); phase B ships the 16 columns to the ranks that finish the rows.
        if (ts == 0) mbar_wait(bar_wout, 0);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = smem_desc_lo(wres_s);
          const uint32_t fa = tmem_base + (uint32_t)(resid & 1) * 128u;
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            umma_tf32_ts(tmem_base + d3_col, fa + ks * 8, make_desc(w_lo + (ks >> 2) * (2048 >> 4) + (ks & 3) * 2), kIdesc16,
                         ks > 0 ? 1u : 0u);
          umma_commit(bar_res_full);
        }
        __syncwarp();
      }
    } else if (warp == kTcMmaWarp + 1) {
      // The spare warp brings the memoised table branches of the rows this rank finishes (lane = row 32 * rank + lane, all 256
      // hidden columns) into the exchange layout [c4][row] while phase A runs, so that phase C reads them as one more
      // coalesced source.  (Threads reading the table rows directly in phase C would touch 32 cache lines per instruction.)
      if (memo) {
        float4 *const xm = A.scratch + (size_t)tile * kScratchF4PerTile + kScratchPartialF4 + (size_t)rank * kScratchBlockF4 + lane;
        const int e = tile * 128 + (int)rank * 32 + lane;
        const float4 *src = reinterpret_cast<const float4 *>(
            A.memo + (size_t)memo_row_of(kFused ? F.S.state : A.memo_state, e < A.n ? e : 0, kFused ? F.S.n_chunks : A.memo_n_chunks) * 256);
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
          float4 v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __ldg(src + pass * 16 + j);
#pragma unroll
          for (int j = 0; j < 16; ++j) __stcg(xm + (size_t)(pass * 16 + j) * 32, v[j]);
        }
      }
    } else if (warp >= kTcEpiWarp0) {
      for (int i = 0; i < nloc; ++i) {
        const uint32_t buf = i & 1u;
        const int bi = P.branch[i];
        mbar_wait(bar_d1_full + 8 * buf, d1_use[buf] & 1u);
        ++d1_use[buf];
        tc_fence_after();
        if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[384 + 2 * i] = clock64();
#pragma unroll 1
        for (int c = 2 * half; c < 2 * half + 2; ++c) {
          float v[32];
          const uint32_t ta = tmem_base + lane_addr + buf * 128u + c * 32u;
          tmem_ld32(ta, v);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = leaky(v[jj] + K.bias1[bi][c * 32 + jj]);
          tmem_st32(ta, v);
          if (!kFused && A.feat_dbg && live) {
            float *dst = A.feat_dbg + (size_t)env * (K.n_branches * kHidden) + bi * kHidden + c * 32;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) dst[jj] = v[jj];
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_feat_full + 8 * buf);
        if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[384 + 2 * i + 1] = clock64();
      }
      mbar_wait(bar_d2_full, par);   // every MMA of this CTA has completed: partial D2 final, TMA stages idle
      tc_fence_after();
    }
    if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[480] = clock64();

    MANSY_DBG(5);
    // ================= phase B: reduce-scatter the partials through L2, split by ROWS =================
    // Rank d finishes rows 32d .. 32d+31 of the tile for all 256 hidden columns (and then their heads: no second exchange).
    // Rows 32q .. 32q+31 are TMEM lane quarter q, which warps q, q+4, q+8, q+12 can read: all sixteen warps take part, warp w
    // moves columns 64 (w >> 2) .. +63 of quarter w & 3 -- to rank q's block of the exchange buffer, or, for this rank's own
    // quarter, to shared memory.  Layout of a block: [c4 = float4 column 0..63][row 0..31] (lanes = rows: coalesced both ways).
    float4 *const xch = A.scratch + (size_t)tile * kScratchF4PerTile;
    {
      if (warp < kTcEpiWarp0) {          // (the epilogue warps waited for this phase at the end of phase A)
        mbar_wait(bar_d2_full, par);     // every MMA of this CTA has completed: partial D2 final, TMA stages idle
        tc_fence_after();
      }
      const uint32_t bq = (uint32_t)warp & 3u, cpair = (uint32_t)warp >> 2;
      const uint32_t laddr = (bq * 32u) << 16;
#pragma unroll 1
      for (uint32_t hh = 0; hh < 2; ++hh) {
        const uint32_t chunk = 2u * cpair + hh;      // 32 columns
        float v[32];
        tmem_ld32(tmem_base + laddr + 256u + chunk * 32u, v);
        if (bq == rank) {
          const uint32_t dst = own_s + ((chunk * 8u) * 32u + (uint32_t)lane) * 16u;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + c4 * 512), "f"(v[4 * c4]), "f"(v[4 * c4 + 1]),
                         "f"(v[4 * c4 + 2]), "f"(v[4 * c4 + 3]) : "memory");
        } else {
          float4 *dst = xch + (size_t)(bq * 4u + rank) * kScratchBlockF4 + (chunk * 8u) * 32u + (uint32_t)lane;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) __stcg(dst + c4 * 32, make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]));
        }
      }
      if (resid >= 0 && cpair == 0) {
        // the residual's 16 head columns of quarter bq's rows -> the rank that finishes them (through L2 for this rank too)
        mbar_wait(bar_res_full, par);
        tc_fence_after();
        float a16[16];
        tmem_ld16(tmem_base + laddr + d3_col, a16);
        float4 *dst = xch + kScratchPartialF4 + kScratchMemoF4 + (size_t)bq * 128 + (uint32_t)lane;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) __stcg(dst + c4 * 32, make_float4(a16[4 * c4], a16[4 * c4 + 1], a16[4 * c4 + 2], a16[4 * c4 + 3]));
      }
      tc_fence_before();
    }
    if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[482] = clock64();
    cluster_sync_all();              // (1) partials of all four ranks are in L2 (release / acquire at cluster scope)
    MANSY_DBG(10);
    if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[483] = clock64();

    // ================= phase C: this rank's 32 rows -- hidden layer, then the heads from shared memory =================
    if (warp >= kTcEpiWarp0 || warp < 4) {
      // 384 threads, item f = c4 * 32 + row (2 048 float4 per source): own partial (shared memory) + the three other ranks'
      // (ascending) + the memoised table branches, bias, LeakyReLU -> the heads' A operand: box c4 >> 3 of [128 rows x 32 floats]
      // in the 128-byte-swizzled K-major layout the TMA boxes have, rows 96 .. 127 (so that the heads land in TMEM lane quarter
      // 3, the spare warp's).
      const uint32_t t12 = (uint32_t)(warp < 4 ? warp : warp - 4) * 32u + (uint32_t)lane;
      const float4 *const xin = xch + (size_t)(rank * 4u) * kScratchBlockF4;
      const float4 *const xmemo = xch + kScratchPartialF4 + (size_t)rank * kScratchBlockF4;
      const uint32_t row_s = 96u + (uint32_t)lane;
#pragma unroll 1
      for (uint32_t base_f = 0; base_f < 2048u; base_f += 3u * 384u) {
        float4 acc[3], sv[4][3];
        uint32_t f[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) f[u] = base_f + (uint32_t)u * 384u + t12;
        // all twelve loads of the pass are in flight together (one L2 round trip per pass, not one per source) ...
#pragma unroll
        for (uint32_t j3 = 0; j3 < 4; ++j3) {
          const float4 *from = j3 < 3 ? xin + (size_t)(j3 + (j3 >= rank ? 1u : 0u)) * kScratchBlockF4 : xmemo;
#pragma unroll
          for (int u = 0; u < 3; ++u)
            sv[j3][u] = (f[u] < 2048u && (j3 < 3 || memo)) ? __ldcg(from + f[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) acc[u] = f[u] < 2048u ? lds128(own_s + f[u] * 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
        // ... and added in a fixed order: own + the other ranks ascending + memo
#pragma unroll
        for (uint32_t j3 = 0; j3 < 4; ++j3) {
          if (j3 == 3 && !memo) break;
#pragma unroll
          for (int u = 0; u < 3; ++u) { acc[u].x += sv[j3][u].x; acc[u].y += sv[j3][u].y; acc[u].z += sv[j3][u].z; acc[u].w += sv[j3][u].w; }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (f[u] < 2048u) {
            const uint32_t c4 = f[u] >> 5;               // (f & 31 == lane: the row)
            float4 h;
            h.x = leaky(acc[u].x + K.bias2[c4 * 4u]); h.y = leaky(acc[u].y + K.bias2[c4 * 4u + 1]);
            h.z = leaky(acc[u].z + K.bias2[c4 * 4u + 2]); h.w = leaky(acc[u].w + K.bias2[c4 * 4u + 3]);
            const uint32_t dst = hid_s + (c4 >> 3) * kBoxBytes + row_s * 128u + (((c4 & 7u) ^ (row_s & 7u)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
            if (!kFused && A.hid_dbg) {      // NOTE: without the residual (it enters through the heads in this kernel)
              const int e = tile * 128 + (int)rank * 32 + lane;
              if (e < A.n) *reinterpret_cast<float4 *>(A.hid_dbg + (size_t)e * 256 + c4 * 4u) = h;
            }
          }
        }
      }
      fence_async_smem();              // the tensor pipe (async proxy) reads what the generic proxy has just stored
      mbar_arrive(bar_hid_full);
    } else if (warp == kTcMmaWarp) {
      constexpr uint32_t kIdesc16 = idesc_tf32(16);
      if (resid < 0 && ts == 0) mbar_wait(bar_wout, 0);
      mbar_wait(bar_hid_full, par);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = smem_desc_lo(hid_s), w_lo = smem_desc_lo(wout_s);
#pragma unroll
        for (int ks = 0; ks < 32; ++ks)      // K = 256 hidden units: 8 boxes of 4 K-steps
          umma_tf32(tmem_base + 256u, make_desc(a_lo + (ks >> 2) * (kBoxBytes >> 4) + (ks & 3) * 2),
                    make_desc(w_lo + (ks >> 2) * (2048 >> 4) + (ks & 3) * 2), kIdesc16, ks > 0 ? 1u : 0u);
        umma_commit(bar_d3_full);
      }
      __syncwarp();
    }
    if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[484] = clock64();
    // (No second cluster barrier: the heads of a rank's rows are computed where the rows are finished.)
    MANSY_DBG(20);
    if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[485] = clock64();

    if (kFused && warp == 0) {
      // Which chunk the NEXT observation describes does not depend on the action about to be sampled
      // (simulator.py:105-106: next_chunk += 1; mansy_env.py:100-101: the sample after a finished episode), so its
      // table columns (next chunk sizes / qualities: floats 8..647 of a MANSY row, 8..327 of a SimpleRL row = one row of
      // obs_tab / size_norm) are copied by ONE warp as bulk copies -- lane = environment: table row -> the idle TMA
      // stages -> observation row -- while the epilogue warps sample and step.  No register staging and nothing in the
      // L1 load/store queue the simulator phase waits on.  (After cluster barrier 2: every MMA that read the stages
      // has completed -- the epilogue warps waited for bar_d2_full before they arrived there.)
      mbar_wait(bar_d3_full, par);       // stages 0 and 1 hold the heads' A operand until the head MMAs have completed
      const int ei = tile * 128 + (int)rank * 32 + lane;
      const int nxt = (int)((t + 1) % F.slabs);
      const bool ok = ei < A.n;
      size_t trow = 0;
      if (ok) {
        EnvState nx;
        load_state(F.S, ei, nx);
        int video, pair, start, chunk;
        next_obs_chunk(F.S, nx, 1, video, pair, start, chunk);
        trow = (size_t)video * F.S.n_chunks + chunk;
      }
      asm volatile("bar.arrive 2, 288;" ::: "memory");     // state(t) has been read: the simulator phase may overwrite it
    MANSY_DBG(21);
      constexpr uint32_t kRowBytes = MODE == MANSY_OBS_MANSY ? 2560u : 1280u;
      const uint32_t grp = (uint32_t)lane >> 3;            // one mbarrier per 8 environments (<= 20 KB of transactions each)
      const uint32_t n_grp = __popc(__ballot_sync(0xFFFFFFFFu, ok) & (0xFFu << (8 * grp)));
      const uint32_t tb = bar_tab + 8 * grp;
      if ((lane & 7) == 0 && n_grp) mbar_expect_tx(tb, n_grp * kRowBytes);
      __syncwarp();
      const uint32_t sdst = stage0 + (uint32_t)lane * kRowBytes;
      if (ok) bulk_load(sdst, MODE == MANSY_OBS_MANSY ? F.S.obs_tab + trow * (2 * kTableRow) : F.S.size_norm + trow * kTableRow, kRowBytes, tb);
      MANSY_DBG(22);
      if (n_grp) { mbar_wait(tb, tab_use & 1u, 77); ++tab_use; }
      MANSY_DBG(23);
      if (ok) {
        bulk_store(F.obs + ((size_t)nxt * A.n + ei) * F.obs_stride + 8, sdst, kRowBytes);
        bulk_commit();
      }
      MANSY_DBG(24);
      bulk_wait_all();       // rows written (a non-memo policy reads them next step) and the stages free for the next TMA loads
    }
    // fused: the simulator phase's action-independent loads (state record, history slot, viewport / trace entries) are issued
    // by all eight epilogue warps now, so they land while warp q == rank samples the actions.  (Measured alternatives that
    // did not pay, profiles/r02*_fused_timeline.txt: sampling spread over the 8 lanes of each environment, the loads staged
    // through cp.async into an idle TMA stage -- each lengthens this phase by as much as it shortens the next.)
    if (kFused && sim_live) {
      load_state(F.S, sim_i, sim_st);
      load_slot(F.S, sim_i, sim_et & 7, sim_slot);
      sim_in = step_prefetch(F.S, sim_st, sim_et & 7);
      sim_pred = __ldg(F.S.vp_pred + (size_t)sim_st.pair * F.S.n_vp_chunks + (min(sim_st.next_chunk + 1, sim_st.end_chunk) - sim_st.start_chunk));
      // What depends on the action about to be sampled -- which of the 16 (pair, chunk, action) outcomes -- and what an
      // episode end needs -- the next sample's record -- go to the idle third TMA stage through cp.async: no destination
      // registers, nothing waits for them before the step itself.
      const uint32_t el = (uint32_t)sim_et >> 3, sb = (uint32_t)sim_et & 7u;
      if (F.S.outcome != nullptr && F.outcome_prefetch) {
        const uint4 *o = reinterpret_cast<const uint4 *>(F.S.outcome + sim_in.vi * kOutcomeActions + 2 * sb);
        const uint32_t dst = sim_s + (el * 8u + sb) * 64u;
#pragma unroll
        for (int j = 0; j < 4; ++j) cp_async16(dst + 16u * j, o + j);
        sim_in.outcomes_smem = sim_s + el * 512u;
      }
      if (sb < 3 && sim_st.next_chunk + 1 > sim_st.end_chunk)
        cp_async16(sim_s + 16384u + el * 64u + sb * 16u, reinterpret_cast<const uint4 *>(F.S.ep_init + sim_st.cursor % F.S.n_samples) + sb);
    }
    if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[493] = clock64();

    // ================= phase D: each rank finishes its 32 rows =================
    // On the spare warp (lane = row 32 * rank + lane of the tile): in the fused kernel the eight epilogue warps are walking
    // the dependent load chain of the simulator phase meanwhile, so the action is in shared memory by the time they need it
    // instead of ~2 200 cycles after their loads were issued.
    if (warp == kTcMmaWarp + 1) {
      const int env = tile * 128 + (int)rank * 32 + lane;
      const bool live = env < A.n;
      float4 res[4];
      if (K.residual_slot >= 0) {          // requested before the wait below: it hides their latency
        const float4 *xr = xch + kScratchPartialF4 + kScratchMemoF4 + (size_t)rank * 128 + lane;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) res[c4] = __ldcg(xr + c4 * 32);
      }
      mbar_wait(bar_d3_full, par);         // heads of rows 32 * rank .. + 31 in TMEM lanes 96 .. 127 (this warp's quarter)
      tc_fence_after();
      float acc[16];
      tmem_ld16(tmem_base + (96u << 16) + 256u, acc);
      tc_fence_before();
      if (K.residual_slot >= 0) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          acc[4 * c4] += res[c4].x; acc[4 * c4 + 1] += res[c4].y; acc[4 * c4 + 2] += res[c4].z; acc[4 * c4 + 3] += res[c4].w;
        }
      }
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[o] += K.bout[o];
      int act = 0;
      if (live) {
        const size_t orow = kFused ? (size_t)cur * A.n + env : (size_t)env;   // fused: outputs of step t live in slab t % slabs
        float p[kActions];
#pragma unroll
        for (int o = 0; o < kActions; ++o) p[o] = acc[o];
        if (K.softmax == 2) {     // QoE identifier: torch.sigmoid on its 3 outputs (mansy.py:141)
#pragma unroll
          for (int o = 0; o < 3; ++o) p[o] = 1.0f / (1.0f + expf(-p[o]));
        } else if (K.softmax) {
          float m = p[0];
#pragma unroll
          for (int o = 1; o < kActions; ++o) m = fmaxf(m, p[o]);
          float sum = 0.f;
#pragma unroll
          for (int o = 0; o < kActions; ++o) { p[o] = expf(p[o] - m); sum += p[o]; }
#pragma unroll
          for (int o = 0; o < kActions; ++o) p[o] = p[o] / sum;
        }
        if (A.value) A.value[orow] = acc[15];
        if (!kFused && A.logits) {
          float4 *dst = reinterpret_cast<float4 *>(A.logits + (size_t)env * 16);
          dst[0] = make_float4(p[0], p[1], p[2], p[3]);
          dst[1] = make_float4(p[4], p[5], p[6], p[7]);
          dst[2] = make_float4(p[8], p[9], p[10], p[11]);
          dst[3] = make_float4(p[12], p[13], p[14], 0.f);
        }
        if (A.actions) {
          float lp;
          categorical_sample(p, K.softmax == 1, A.seed, (uint64_t)(A.env_offset + env), (uint64_t)t, act, lp);
          A.actions[orow] = act;
          if (A.logp) A.logp[orow] = lp;
        }
      }
      if (kFused) {
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(act_s + (uint32_t)lane * 4u), "r"(act) : "memory");
        asm volatile("bar.arrive 1, 288;" ::: "memory");     // the simulator phase picks the actions up from act_s
      }
      if (tl && k == tlk && lane == 0) A.timeline[486] = clock64();
    }

    if (kFused) {
      // ================= phase E: simulator chunk-step of this CTA's 32 environments =================
      if (warp >= kTcEpiWarp0) {
        const int i = sim_i;
        const int sub = sim_et & 7;
        const SimDev &S = F.S;
        EnvState &st = sim_st;
        float (&slot)[8] = sim_slot;
        const bool live_e = sim_live;
        cp_async_wait_all();                                // this lane's share of the staged outcomes / next-sample record ...
        asm volatile("bar.sync 1, 288;" ::: "memory");      // ... and everyone's; the spare warp has put the actions in act_s
        if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[490] = clock64();
        int action;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(action) : "r"(act_s + (uint32_t)(sim_et >> 3) * 4u) : "memory");
        const int nxt = (int)((t + 1) % F.slabs);
        // full-mask collectives when all four environments of the warp step together (always, except in a tail tile)
        auto sim_step = [&](unsigned mask) {
          float reward_f = 0.f;
          bool over = true;
          uint64_t pred = sim_pred;
          if (!(st.flags & kFlagFinished)) {
            const int slot_before = st.ep_step & 7;
            const double reward = step_env(S, st, slot, sub, mask, action, over, nullptr, nullptr, sim_in);
            reward_f = (float)reward;
            if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[491] = clock64() + (reward_f > 1e30f);
            if (sub == slot_before) store_slot(S, i, sub, slot);
            if (over) {
              finish_episode(S, i, st, sub);
              union { EpisodeInit e; float4 q[3]; } u;      // staged before the action was known (the episode end was not in doubt)
              const uint32_t src = sim_s + 16384u + ((uint32_t)sim_et >> 3) * 64u;
              u.q[0] = lds128(src); u.q[1] = lds128(src + 16u); u.q[2] = lds128(src + 32u);
              reset_episode(S, st, u.e);
              pred = u.e.first_pred;
            }
          } else {
            pred = __ldg(S.vp_pred + (size_t)st.pair * S.n_vp_chunks + (min(st.next_chunk, st.end_chunk) - st.start_chunk));
          }
          if (sub == 0) {
            F.reward[(size_t)cur * A.n + i] = reward_f;
            F.done[(size_t)cur * A.n + i] = over ? 1 : 0;
          }
          float *row = F.obs + ((size_t)nxt * A.n + i) * F.obs_stride;
          emit_obs_pred<MODE>(pred, sub, row);
          emit_obs_dynamic<MODE>(S, st, slot, sub, mask, row);
        };
        if (__all_sync(0xFFFFFFFFu, live_e && !(st.flags & kFlagFinished))) sim_step(0xFFFFFFFFu);
        else if (live_e) sim_step(group_mask());
        MANSY_DBG(32);
        asm volatile("bar.sync 2, 288;" ::: "memory");       // the table-row warp has read state(t)
        MANSY_DBG(33);
        if (i < A.n) store_state(S, i, st, sub);
        if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[492] = clock64();
        // the next step's TMA (async proxy) reads the rows just written through the generic proxy
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[487] = clock64();
      MANSY_DBG(39);
      cluster_sync_all();            // (3) the next observation rows of the tile are complete
      MANSY_DBG(40);
      // (No proxy fence on the consumer side: every writer of the rows ran fence.proxy.async before it arrived at the
      // barrier, which is what orders its generic-proxy stores before the TMA loads issued after the barrier.  A second
      // fence in the producer warps delayed the first load of every step by ~700 cycles.)
      if (tl && k == tlk && threadIdx.x == 32 * kTcEpiWarp0) A.timeline[488] = clock64();
    }
   }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D fp32 row-major tensor [rows][cols] with `row_stride` floats between rows; box = [box_rows][32 floats],
// SWIZZLE_128B, out-of-bounds elements read as zero.
int make_map(CUtensorMap *m, const float *ptr, uint64_t cols, uint64_t rows, uint64_t row_stride, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(MANSY_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MANSY_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return MANSY_OK;
}

std::mutex g_slot_mutex;
bool g_slot_used[kTcSlots] = {false, false, false};

struct BranchPlan {
  int canon;       // index in FeatureNet concat order
  int off, k;      // observation floats [off, off + k)
  bool extra;      // weights live in the extra W1 box (the branch shares its K-step with another one)
};

}  // namespace

int tc_make_map(CUtensorMap *m, const float *ptr, uint64_t cols, uint64_t rows, uint64_t row_stride, uint32_t box_rows) {
  return make_map(m, ptr, cols, rows, row_stride, box_rows);
}

int tc_create(mansy_policy *p, const mansy_policy_weights_t *w) {
  const bool is_ident = w->kind == MANSY_NET_IDENTIFIER;
  const bool is_mansy = w->kind == MANSY_OBS_MANSY || is_ident;       // MANSY observation rows
  const int nb = is_mansy ? 10 : 5;
  const int n_out = is_ident ? 3 : kActions;
  const int obs_floats = is_mansy ? MANSY_OBS_MANSY_STRIDE : MANSY_OBS_SIMPLE_STRIDE;
  // processing order: big branches first, the residual branch last (its feature tile must still
  // be in shared memory when the heads run); `extra` marks the branch whose K-step is shared.
  static const BranchPlan mansy_plan[10] = {
      {1, 8, 320, false},  {2, 328, 320, false}, {3, 648, 64, false}, {0, 0, 8, false},   {4, 728, 8, false},
      {5, 736, 8, false},  {6, 744, 8, false},   {7, 752, 8, false},  {8, 779, 1, true},  {9, 776, 3, false}};
  static const BranchPlan simple_plan[5] = {
      {1, 8, 320, false}, {4, 328, 64, false}, {0, 0, 8, false}, {2, 394, 1, true}, {3, 392, 2, false}};
  // QoE identifier: fc2 reads the 15 action_one_hot floats 760..774 (K-steps 95 and 96: boxes 23 and 24)
  static const BranchPlan ident_plan[10] = {
      {1, 8, 320, false},  {2, 328, 320, false}, {3, 648, 64, false}, {0, 0, 8, false},   {4, 728, 8, false},
      {5, 736, 8, false},  {6, 744, 8, false},   {7, 752, 8, false},  {8, 779, 1, true},  {9, 760, 15, false}};
  const BranchPlan *plan = is_ident ? ident_plan : (is_mansy ? mansy_plan : simple_plan);
  const int main_boxes = (obs_floats + 31) / 32;          // 25 / 13
  const int w1_cols = (main_boxes + 1) * 32;
  const int F = nb * kHidden;

  TcState *st = new (std::nothrow) TcState();
  if (!st) return set_error(MANSY_E_NOMEM, "out of host memory");
  st->obs_floats = obs_floats;
  st->n_branches = nb;
  {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    for (int s = 0; s < kTcSlots; ++s)
      if (!g_slot_used[s]) { g_slot_used[s] = true; st->slot = s; break; }
  }
  if (st->slot < 0) {
    delete st;
    return set_error(MANSY_E_STATE, "all tensor-core policy slots are in use (destroy another policy first)");
  }

  TcConst *hc = new (std::nothrow) TcConst();
  if (!hc) {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    g_slot_used[st->slot] = false;
    delete st;
    return set_error(MANSY_E_NOMEM, "out of host memory");
  }
  memset(hc, 0, sizeof(*hc));
  std::vector<float> w1img((size_t)kHidden * w1_cols, 0.f), wfcimg((size_t)256 * F, 0.f);
  for (int i = 0; i < nb; ++i) {
    const BranchPlan &b = plan[i];
    for (int f = 0; f < kHidden; ++f) {
      for (int k = 0; k < b.k; ++k) {
        const int col = b.extra ? main_boxes * 32 + ((b.off + k) % 32) : b.off + k;
        w1img[(size_t)f * w1_cols + col] = w->branch_w[b.canon][(size_t)f * b.k + k];
      }
      hc->bias1[i][f] = w->branch_b[b.canon][f];
      for (int j = 0; j < kHidden; ++j) {
        wfcimg[(size_t)j * F + i * kHidden + f] = w->actor_fc_w[(size_t)j * F + b.canon * kHidden + f];
        if (!is_ident)
          wfcimg[(size_t)(kHidden + j) * F + i * kHidden + f] = w->critic_fc_w[(size_t)j * F + b.canon * kHidden + f];
      }
    }
  }
  std::vector<float> woutimg((size_t)16 * 256, 0.f);   // rows 0..14: [actor.out | 0], row 15: [0 | critic.out]
  for (int j = 0; j < kHidden; ++j) {
    for (int o = 0; o < n_out; ++o) woutimg[(size_t)o * 256 + j] = w->actor_out_w[(size_t)o * kHidden + j];
    if (!is_ident) woutimg[(size_t)15 * 256 + kHidden + j] = w->critic_out_w[j];
    hc->bias2[j] = w->actor_fc_b[j];
    hc->bias2[kHidden + j] = is_ident ? 0.f : w->critic_fc_b[j];
  }
  std::vector<float> wresimg((size_t)16 * kHidden, 0.f);   // residual through the heads: rows 0..14 actor.out, row 15 critic.out
  for (int j = 0; j < kHidden; ++j) {
    for (int o = 0; o < n_out; ++o) wresimg[(size_t)o * kHidden + j] = w->actor_out_w[(size_t)o * kHidden + j];
    if (!is_ident) wresimg[(size_t)15 * kHidden + j] = w->critic_out_w[j];
  }
  for (int o = 0; o < n_out; ++o) hc->bout[o] = w->actor_out_b[o];
  if (!is_ident) hc->bout[15] = w->critic_out_b[0];
  hc->n_branches = nb;
  hc->softmax = is_ident ? 2 : (is_mansy ? 0 : 1);     // output activation: none / softmax / sigmoid on 3 outputs
  hc->residual_slot = is_mansy ? nb - 1 : -1;

  // job list: L1(p0), L1(p1), L2(p0), L1(p2), L2(p1), ..., L1(p_last), L2(p_last-1), L2(p_last)
  int nj = 0;
  const int pair1 = getenv("MANSY_TC_PAIR1") ? atoi(getenv("MANSY_TC_PAIR1")) : 2;   // debugging knobs: boxes / chunks per job
  const int pair2 = getenv("MANSY_TC_PAIR2") ? atoi(getenv("MANSY_TC_PAIR2")) : 2;
  // A TMA box may start at any 16-byte aligned column: a branch's boxes start at the branch's first float (all offsets are
  // multiples of 8 floats = one K-step), not at the multiple of 32 floats below it -- the 64 floats at 648 are 2 boxes = one
  // job instead of 3 boxes = two jobs, a 320-float branch 10 boxes instead of 11.  The K-steps and their order are the same
  // either way (bit-identical sums).  The `extra` branch keeps the 32-aligned box its weight box mirrors.
  const bool aligned = !(getenv("MANSY_TC_BRANCH_BOXES") && getenv("MANSY_TC_BRANCH_BOXES")[0] == '0');
  auto push_l1 = [&](int i) {
    const BranchPlan &b = plan[i];
    const int origin = (aligned && !b.extra && b.off % 8 == 0) ? b.off / 8 : (b.off / 32) * 4;   // K-step the first box starts at
    const int lo = b.off / 8 - origin, hi = (b.off + b.k + 7) / 8 - origin;      // K-steps of 8 floats, relative to it
    const int box_lo = lo / 4, box_hi = (hi - 1) / 4;          // boxes of 4 K-steps; a job takes up to two
    for (int box = box_lo; box <= box_hi; box += pair1) {
      const int nbox = (pair1 == 2 && box + 1 <= box_hi) ? 2 : 1;
      TcJob &j = hc->jobs[nj++];
      j.type = kJobL1; j.a_box = (uint8_t)(origin * 2 + box * 8); j.w_box = (uint8_t)(b.extra ? main_boxes * 8 : origin * 2 + box * 8);
      j.s_lo = (uint8_t)((lo > box * 4 ? lo : box * 4) - box * 4);
      j.s_hi = (uint8_t)((hi < (box + nbox) * 4 ? hi : (box + nbox) * 4) - box * 4);
      j.slot = (uint8_t)i; j.chunk = (uint8_t)nbox;
      j.flags = (uint8_t)((box == box_lo ? kFlagFirst : 0) | (box + nbox > box_hi ? kFlagLast : 0));
    }
  };
  auto push_l2 = [&](int i) {
    for (int c = 0; c < 4; c += pair2) {
      TcJob &j = hc->jobs[nj++];
      j.type = kJobL2; j.a_box = j.w_box = j.s_lo = 0; j.s_hi = (uint8_t)pair2; j.slot = (uint8_t)i; j.chunk = (uint8_t)c;
      j.flags = (uint8_t)((c + pair2 == 4 ? kFlagLast : 0) | (i == 0 && c == 0 ? kFlagTileFirstL2 : 0) |
                          (i == nb - 1 && c + pair2 == 4 ? kFlagTileLastL2 : 0));
    }
  };
  auto branch_boxes = [&](const BranchPlan &b) {
    const int origin = (aligned && !b.extra && b.off % 8 == 0) ? b.off / 8 : (b.off / 32) * 4;
    return ((b.off + b.k + 7) / 8 - origin - 1) / 4 + 1;
  };
  push_l1(0);
  for (int i = 1; i < nb; ++i) { push_l1(i); push_l2(i - 1); }
  push_l2(nb - 1);
  hc->n_jobs = nj;

  // split-K plan: longest-processing-time assignment of the branches to the 4 cluster ranks by operand bytes
  // (layer-1 boxes of observation rows + W1, 128 KB of Wfc rows); the plan is already sorted by size and the
  // residual branch comes last, so it ends up last on its rank (its features must stay in tensor memory).
  {
    int load[kTcRanks] = {0, 0, 0, 0};
    for (int r = 0; r < kTcRanks; ++r) { hc->rank[r].n_branches = 0; hc->rank[r].resid_local = -1; }
    for (int i = 0; i < nb; ++i) {
      const BranchPlan &b = plan[i];
      const int boxes = branch_boxes(b);
      int best = 0;
      for (int r = 1; r < kTcRanks; ++r) if (load[r] < load[best]) best = r;
      load[best] += boxes * 32 + 128;
      TcRankPlan &rp = hc->rank[best];
      if (i == hc->residual_slot) rp.resid_local = rp.n_branches;
      rp.branch[rp.n_branches++] = (uint8_t)i;
    }
    for (int r = 0; r < kTcRanks; ++r) {
      TcRankPlan &rp = hc->rank[r];
      const int nl = rp.n_branches;
      int cnt = 0;             // push_l1 / push_l2 append to hc->jobs: use its tail as scratch and copy out
      auto emit_l1 = [&](int li) {
        const int before = nj;
        push_l1(rp.branch[li]);
        for (int j = before; j < nj; ++j) { rp.jobs[cnt] = hc->jobs[j]; rp.jobs[cnt].slot = (uint8_t)li; ++cnt; }
        nj = before;
      };
      auto emit_l2 = [&](int li) {
        const int before = nj;
        push_l2(rp.branch[li]);
        for (int j = before; j < nj; ++j) {
          TcJob t = hc->jobs[j];
          t.slot = (uint8_t)li;
          t.flags = (uint8_t)(t.flags & ~(kFlagTileFirstL2 | kFlagTileLastL2));
          if (li == 0 && t.chunk == 0) t.flags |= kFlagTileFirstL2;
          if (li == nl - 1 && (t.flags & kFlagLast)) t.flags |= kFlagTileLastL2;
          rp.jobs[cnt++] = t;
        }
        nj = before;
      };
      if (nl > 0) {
        emit_l1(0);
        for (int li = 1; li < nl; ++li) { emit_l1(li); emit_l2(li - 1); }
        emit_l2(nl - 1);
      }
      rp.n_jobs = cnt;
    }
    // the same plans without the memoised 320-input branches (k == 320: next chunk sizes / qualities)
    auto build_plan = [&](TcRankPlan &rp) {      // rp.branch[0 .. n_branches) filled; emits its job list
      const int nl = rp.n_branches;
      int cnt = 0;
      auto emit_l1 = [&](int li) {
        const int before = nj;
        push_l1(rp.branch[li]);
        for (int j = before; j < nj; ++j) { rp.jobs[cnt] = hc->jobs[j]; rp.jobs[cnt].slot = (uint8_t)li; ++cnt; }
        nj = before;
      };
      auto emit_l2 = [&](int li) {
        const int before = nj;
        push_l2(rp.branch[li]);
        for (int j = before; j < nj; ++j) {
          TcJob t = hc->jobs[j];
          t.slot = (uint8_t)li;
          t.flags = (uint8_t)(t.flags & ~(kFlagTileFirstL2 | kFlagTileLastL2));
          if (li == 0 && t.chunk == 0) t.flags |= kFlagTileFirstL2;
          if (li == nl - 1 && (t.flags & kFlagLast)) t.flags |= kFlagTileLastL2;
          rp.jobs[cnt++] = t;
        }
        nj = before;
      };
      if (nl > 0) {
        emit_l1(0);
        for (int li = 1; li < nl; ++li) { emit_l1(li); emit_l2(li - 1); }
        emit_l2(nl - 1);
      }
      rp.n_jobs = cnt;
    };
    {
      int mload[kTcRanks] = {0, 0, 0, 0};
      for (int r = 0; r < kTcRanks; ++r) { hc->rank_memo[r].n_branches = 0; hc->rank_memo[r].resid_local = -1; }
      hc->solo_memo.n_branches = 0; hc->solo_memo.resid_local = -1;
      for (int i = 0; i < nb; ++i) {
        const BranchPlan &b = plan[i];
        if (b.k == 320) continue;
        const int boxes = branch_boxes(b);
        int best = 0;
        for (int r = 1; r < kTcRanks; ++r) if (mload[r] < mload[best]) best = r;
        mload[best] += boxes * 32 + 128;
        TcRankPlan &rp = hc->rank_memo[best];
        if (i == hc->residual_slot) rp.resid_local = rp.n_branches;
        rp.branch[rp.n_branches++] = (uint8_t)i;
        if (i == hc->residual_slot) hc->solo_memo.resid_local = hc->solo_memo.n_branches;
        hc->solo_memo.branch[hc->solo_memo.n_branches++] = (uint8_t)i;
      }
      for (int r = 0; r < kTcRanks; ++r) build_plan(hc->rank_memo[r]);
      build_plan(hc->solo_memo);
    }
    memset(&hc->jobs[nj], 0, sizeof(TcJob) * (kTcMaxJobs - nj));     // scratch entries used above
  }

  int rc = MANSY_OK;
  float *d_w1 = nullptr, *d_wfc = nullptr, *d_wout = nullptr, *d_wres = nullptr;
  if (cudaMalloc(&d_w1, w1img.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_wfc, wfcimg.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_wout, woutimg.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_wres, wresimg.size() * sizeof(float)) != cudaSuccess)
    rc = set_error(MANSY_E_NOMEM, "cudaMalloc failed (tensor-core weight images)");
  if (d_w1) p->allocs.push_back(d_w1);
  if (d_wfc) p->allocs.push_back(d_wfc);
  if (d_wout) p->allocs.push_back(d_wout);
  if (d_wres) p->allocs.push_back(d_wres);
  if (!rc && (cudaMemcpy(d_w1, w1img.data(), w1img.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpy(d_wfc, wfcimg.data(), wfcimg.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpy(d_wout, woutimg.data(), woutimg.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpy(d_wres, wresimg.data(), wresimg.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpyToSymbol(c_tc, hc, sizeof(TcConst), (size_t)st->slot * sizeof(TcConst)) != cudaSuccess))
    rc = set_error(MANSY_E_CUDA, "upload of the tensor-core weight images failed");
  delete hc;
  if (!rc) rc = make_map(&st->map_w1, d_w1, (uint64_t)w1_cols, kHidden, (uint64_t)w1_cols, 128);
  if (!rc) rc = make_map(&st->map_wfc, d_wfc, (uint64_t)F, 256, (uint64_t)F, 256);
  if (!rc) rc = make_map(&st->map_wout, d_wout, 256, 16, 256, 16);
  if (!rc) rc = make_map(&st->map_wres, d_wres, kHidden, 16, kHidden, 16);
  if (rc) {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    g_slot_used[st->slot] = false;
    delete st;
    return rc;
  }
  p->tc = st;
  return MANSY_OK;
}

// Tensor map of an observation tensor [rows][obs_floats] (row stride `stride` floats), cached per policy.
static int obs_map_for(mansy_policy *p, const float *ptr, uint64_t rows, uint64_t stride, const CUtensorMap **out) {
  TcState *st = p->tc;
  for (auto &m : st->obs_maps)
    if (m.ptr == ptr && m.rows == rows && m.stride == stride) { *out = &m.map; return MANSY_OK; }
  TcState::ObsMap &m = st->obs_maps[st->obs_map_next];
  st->obs_map_next = (st->obs_map_next + 1) % 4;
  m.ptr = nullptr;
  int rc = make_map(&m.map, ptr, (uint64_t)st->obs_floats, rows, stride, 128);
  if (rc) return rc;
  m.ptr = ptr; m.rows = rows; m.stride = stride;
  *out = &m.map;
  return MANSY_OK;
}

// Exchange buffer of the cluster kernel: (re)allocated when a larger batch shows up (not stream-ordered: callers
// that time launches warm up first).
static int ensure_scratch(mansy_policy *p, int n_tiles) {
  TcState *st = p->tc;
  if (st->scratch_tiles >= n_tiles) return MANSY_OK;
  if (st->scratch) { cudaDeviceSynchronize(); cudaFree(st->scratch); st->scratch = nullptr; st->scratch_tiles = 0; }
  void *d = nullptr;
  if (cudaMalloc(&d, (size_t)n_tiles * kScratchF4PerTile * sizeof(float4)) != cudaSuccess)
    return set_error(MANSY_E_NOMEM, "cudaMalloc failed (split-K exchange buffer)");
  st->scratch = static_cast<float4 *>(d);
  st->scratch_tiles = n_tiles;
  return MANSY_OK;
}

void tc_destroy(mansy_policy *p) {
  if (!p || !p->tc) return;
  if (p->tc->scratch) cudaFree(p->tc->scratch);
  {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    g_slot_used[p->tc->slot] = false;
  }
  delete p->tc;
  p->tc = nullptr;
}

}  // namespace mansy

using namespace mansy;

namespace mansy {

// Launches the tensor-core forward (one CTA per tile, or the split-K cluster kernel for small batches).  `pdl`
// adds the programmatic-dependent-launch attribute: the kernel may be scheduled while the previous kernel of
// the stream is still running and orders itself with griddepcontrol.wait before it touches observation rows.
int policy_forward_tc_launch(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                             float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                             int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, int64_t *timeline_dev,
                             int32_t timeline_cta, bool pdl, void *stream, const SimDev *memo_sim) {
  if (!p || !obs_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (!p->tc) return set_error(MANSY_E_STATE, std::string("tensor-core state unavailable: ") + mansy_last_error());
  if (n < 0) return set_error(MANSY_E_INVALID, "n < 0");
  DeviceScope dscope(p->device);
  if (dscope.err != cudaSuccess) return set_error(MANSY_E_CUDA, "cudaSetDevice failed");
  if (obs_stride < p->tc->obs_floats || (obs_stride & 3))
    return set_error(MANSY_E_INVALID, "obs_stride must be >= the padded row length and a multiple of 4 floats");
  if (reinterpret_cast<uintptr_t>(obs_dev) & 15) return set_error(MANSY_E_INVALID, "obs must be 16-byte aligned");
  if (logits_dev && (reinterpret_cast<uintptr_t>(logits_dev) & 15)) return set_error(MANSY_E_INVALID, "logits must be 16-byte aligned");
  if (n == 0) return MANSY_OK;
  const CUtensorMap *map_obs_p = nullptr;
  int rc = obs_map_for(p, obs_dev, (uint64_t)n, (uint64_t)obs_stride, &map_obs_p);
  if (rc) return rc;
  const CUtensorMap &map_obs = *map_obs_p;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.n_tiles = (n + 127) / 128;
  a.logits = logits_dev; a.value = value_dev; a.actions = actions_dev; a.logp = logp_dev;
  a.feat_dbg = feat_dbg_dev; a.hid_dbg = hid_dbg_dev;
  a.seed = seed; a.step = step; a.env_offset = env_offset;
  a.timeline = reinterpret_cast<long long *>(timeline_dev);
  a.timeline_cta = timeline_cta;
  if (memo_sim && !feat_dbg_dev && !hid_dbg_dev) {
    if (n > memo_sim->n_envs) return set_error(MANSY_E_INVALID, "more observation rows than environments in the memo simulator");
    if ((rc = policy_memo_for(p, *memo_sim, stream, &a.memo))) return rc;
    a.memo_state = memo_sim->state;
    a.memo_n_chunks = memo_sim->n_chunks;
  }
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm < 1) n_sm = 148;
  }
  // split-K clusters pay off while every tile's 4-CTA cluster is resident at once (a cluster finishes a tile ~3x sooner
  // but occupies four SMs): up to 37 tiles on 148 SMs.  Beyond that a second wave of clusters costs more than one CTA per
  // tile (measured, tools/rollout_sweep.py: 8 192 envs 38.5 vs 37.1 us per rollout step, 12 288 envs 56.0 vs 40.8 us).
  const bool split4 = p->tc->split == 4 || (p->tc->split == 0 && 4 * a.n_tiles <= n_sm);
  if (split4) {
    if ((rc = ensure_scratch(p, a.n_tiles))) return rc;
    a.scratch = p->tc->scratch;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  static const FusedArgs *fz = new FusedArgs();      // value-initialised: the policy-only instantiation ignores it
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = kTcSmemBytes;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
#define MANSY_TC_LAUNCH(SLOT)                                                                                       \
  do {                                                                                                              \
    static bool attr_done_dev[64] = {false};      /* function attributes are per device */                           \
    bool &attr_done = attr_done_dev[p->device & 63];                                                                \
    if (!attr_done) {                                                                                               \
      e = cudaFuncSetAttribute(policy_tc_kernel<SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes); \
      if (e == cudaSuccess)                                                                                         \
        e = cudaFuncSetAttribute(policy_tc4_kernel<SLOT, MANSY_OBS_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)kTcSmemBytes);                                                                \
      attr_done = (e == cudaSuccess);                                                                               \
    }                                                                                                               \
    if (e == cudaSuccess) {                                                                                         \
      if (split4) {                                                                                                 \
        cfg.gridDim = dim3((unsigned)(kTcRanks * a.n_tiles), 1, 1);                                                 \
        e = cudaLaunchKernelEx(&cfg, policy_tc4_kernel<SLOT, MANSY_OBS_NONE>, map_obs, p->tc->map_w1, p->tc->map_wfc,   \
                               p->tc->map_wout, p->tc->map_wres, a, *fz);                                           \
      } else {                                                                                                      \
        cfg.gridDim = dim3((unsigned)(a.n_tiles < n_sm ? a.n_tiles : n_sm), 1, 1);                                  \
        e = cudaLaunchKernelEx(&cfg, policy_tc_kernel<SLOT>, map_obs, p->tc->map_w1, p->tc->map_wfc, p->tc->map_wout, a); \
      }                                                                                                             \
    }                                                                                                               \
  } while (0)
  switch (p->tc->slot) {
    case 0: MANSY_TC_LAUNCH(0); break;
    case 1: MANSY_TC_LAUNCH(1); break;
    default: MANSY_TC_LAUNCH(2); break;
  }
#undef MANSY_TC_LAUNCH
  if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("policy_tc_kernel launch: ") + cudaGetErrorString(e));
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("policy_tc_kernel launch: ") + cudaGetErrorString(e));
  return MANSY_OK;
}

// Fused rollout: n_steps x (policy forward + sample + simulator step) in ONE launch of the split-K cluster kernel.
// Returns MANSY_OK with *launched = 0 when the configuration does not qualify (the caller then runs the
// two-kernel loop): every cluster must be resident at once for the launch to make progress on all tiles together.
static long long *g_fused_timeline = nullptr;    // debug hook: mansy_debug_fused_timeline
static int g_fused_timeline_cta = 0;

int rollout_fused_launch(mansy_policy *p, const SimDev &S, const mansy_rollout_t *b, int32_t n_steps, int64_t t0,
                         uint64_t seed, void *stream, int *launched) {
  *launched = 0;
  if (!p || !p->tc || !b) return MANSY_OK;
  if (S.obs_mode != MANSY_OBS_MANSY && S.obs_mode != MANSY_OBS_SIMPLE) return MANSY_OK;
  if (p->dev.kind != S.obs_mode || p->tc->split == 1 || n_steps < 1) return MANSY_OK;   // (identifier nets never match)
  const int n = S.n_envs, n_tiles = (n + 127) / 128;
  if (b->obs_stride < p->tc->obs_floats || (b->obs_stride & 3) || (reinterpret_cast<uintptr_t>(b->obs) & 15) ||
      (reinterpret_cast<uintptr_t>(b->logits) & 15))
    return MANSY_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(kTcRanks * n_tiles), 1, 1);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = kTcSmemBytes;
  cfg.stream = s;
  // (the cluster shape is compiled into the kernel: __cluster_dims__)

  const CUtensorMap *map_obs_p = nullptr;     // all slabs as one row-major tensor: row = slab * n + env
  int rc = obs_map_for(p, b->obs, (uint64_t)n * (uint64_t)b->slabs, (uint64_t)b->obs_stride, &map_obs_p);
  if (rc) return rc;
  const CUtensorMap &map_obs = *map_obs_p;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.n_tiles = n_tiles;
  a.logits = b->logits; a.value = b->value; a.actions = b->actions; a.logp = b->logp;
  a.seed = seed; a.step = t0; a.env_offset = S.env_offset;
  a.timeline = g_fused_timeline; a.timeline_cta = g_fused_timeline_cta;
  if ((rc = ensure_scratch(p, n_tiles))) return rc;
  a.scratch = p->tc->scratch;
  if (!getenv("MANSY_NO_POLICY_MEMO")) {
    if ((rc = policy_memo_for(p, S, stream, &a.memo))) return rc;
    a.memo_state = S.state;
    a.memo_n_chunks = S.n_chunks;
  }
  FusedArgs f;
  memset(&f, 0, sizeof(f));
  f.S = S;
  f.obs = b->obs; f.obs_stride = b->obs_stride; f.slabs = b->slabs; f.n_steps = n_steps; f.t0 = t0;
  {
    const char *op = getenv("MANSY_FUSED_OUTCOME_PREFETCH");
    f.outcome_prefetch = (op && op[0] == '0') ? 0 : 1;
  }
  f.actions = b->actions; f.logp = b->logp; f.value = b->value; f.reward = b->reward; f.done = b->done;
  cudaError_t e = cudaSuccess;
#define MANSY_FUSED_LAUNCH(SLOT, MODE)                                                                              \
  do {                                                                                                              \
    static int max_clusters_dev[64];              /* function attributes / occupancy are per device */               \
    static bool max_clusters_init = false;                                                                          \
    if (!max_clusters_init) { for (int i = 0; i < 64; ++i) max_clusters_dev[i] = -1; max_clusters_init = true; }     \
    int &max_clusters = max_clusters_dev[p->device & 63];                                                           \
    if (max_clusters < 0) {                                                                                         \
      e = cudaFuncSetAttribute(policy_tc4_kernel<SLOT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes); \
      int mc = 0;                                                                                                   \
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&mc, policy_tc4_kernel<SLOT, MODE>, &cfg);           \
      if (e == cudaSuccess) max_clusters = mc;                                                                      \
    }                                                                                                               \
    /* One tile per resident cluster is the fast configuration (<= 33 tiles = 4 224 environments on B200: 17.5 us per    \
       step).  Up to TWO tiles per cluster (<= 8 448 environments) the cluster walking its tiles back to back still      \
       beats the two-kernel PDL loop (8 192 envs: 34.8 vs 36.8 us per step, tools/rollout_sweep.py); from three on       \
       the two kernels, each filling the machine, win (12 288 envs: 53.6 vs 40.6 us).  MANSY_FUSED_MULTI_TILE=1 / =0     \
       forces / forbids more than one tile per cluster (tests, A/B runs). */                                              \
    const char *mt = getenv("MANSY_FUSED_MULTI_TILE");                                                               \
    const int tiles_per_cluster = max_clusters >= 1 ? (n_tiles + max_clusters - 1) / max_clusters : 0;               \
    const bool take = tiles_per_cluster == 1 || (mt ? mt[0] != '0' : tiles_per_cluster == 2);                        \
    if (e == cudaSuccess && max_clusters >= 1 && take) {                                                             \
      cfg.gridDim = dim3((unsigned)(kTcRanks * (n_tiles < max_clusters ? n_tiles : max_clusters)), 1, 1);           \
      e = cudaLaunchKernelEx(&cfg, policy_tc4_kernel<SLOT, MODE>, map_obs, p->tc->map_w1, p->tc->map_wfc, p->tc->map_wout, \
                             p->tc->map_wres, a, f);                                                                \
      if (e == cudaSuccess) *launched = 1;                                                                          \
    }                                                                                                               \
  } while (0)
  if (S.obs_mode == MANSY_OBS_MANSY) {
    switch (p->tc->slot) {
      case 0: MANSY_FUSED_LAUNCH(0, MANSY_OBS_MANSY); break;
      case 1: MANSY_FUSED_LAUNCH(1, MANSY_OBS_MANSY); break;
      default: MANSY_FUSED_LAUNCH(2, MANSY_OBS_MANSY); break;
    }
  } else {
    switch (p->tc->slot) {
      case 0: MANSY_FUSED_LAUNCH(0, MANSY_OBS_SIMPLE); break;
      case 1: MANSY_FUSED_LAUNCH(1, MANSY_OBS_SIMPLE); break;
      default: MANSY_FUSED_LAUNCH(2, MANSY_OBS_SIMPLE); break;
    }
  }
#undef MANSY_FUSED_LAUNCH
  if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("fused rollout kernel: ") + cudaGetErrorString(e));
  if (*launched) {
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("fused rollout kernel launch: ") + cudaGetErrorString(e));
  }
  return MANSY_OK;
}

}  // namespace mansy

extern "C" {

int mansy_policy_forward_tc(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                            float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                            int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, void *stream) {
  return policy_forward_tc_launch(p, obs_dev, obs_stride, n, logits_dev, value_dev, actions_dev, logp_dev, seed, step,
                                  env_offset, feat_dbg_dev, hid_dbg_dev, nullptr, 0, false, stream);
}

int mansy_policy_forward_tc_sim(mansy_policy_t p, mansy_handle_t h, const float *obs_dev, int64_t obs_stride, float *logits_dev,
                                float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step, void *stream) {
  if (!h) return set_error(MANSY_E_INVALID, "NULL argument");
  const SimDev *S = sim_dev_of(h);
  return policy_forward_tc_launch(p, obs_dev, obs_stride, S->n_envs, logits_dev, value_dev, actions_dev, logp_dev, seed, step,
                                  S->env_offset, nullptr, nullptr, nullptr, 0, false, stream, getenv("MANSY_NO_POLICY_MEMO") ? nullptr : S);
}

int mansy_policy_forward_tc_timeline(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                                     float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                                     int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, int64_t *timeline_dev,
                                     void *stream) {
  const char *cta = getenv("MANSY_TC_TIMELINE_CTA");
  return policy_forward_tc_launch(p, obs_dev, obs_stride, n, logits_dev, value_dev, actions_dev, logp_dev, seed, step,
                                  env_offset, feat_dbg_dev, hid_dbg_dev, timeline_dev, cta ? atoi(cta) : 0, false, stream);
}

int mansy_debug_fused_timeline(int64_t *timeline_dev, int32_t cta) {
  g_fused_timeline = reinterpret_cast<long long *>(timeline_dev);
  g_fused_timeline_cta = cta;
  if (const char *st = getenv("MANSY_TC_TIMELINE_STEP")) g_fused_timeline_cta |= (atoi(st) + 1) << 16;
#ifdef MANSY_STEP_PROFILE
  long long *sp = timeline_dev ? reinterpret_cast<long long *>(timeline_dev) + 496 : nullptr;   // stamps [496..501]
  cudaMemcpyToSymbol(d_step_prof, &sp, sizeof(sp));
#endif
  return MANSY_OK;
}

int mansy_debug_progress(int32_t *progress_dev) {
#ifdef MANSY_MBAR_WATCHDOG
  volatile int *ptr = progress_dev;
  if (cudaMemcpyToSymbol(d_dbg_progress, &ptr, sizeof(ptr)) != cudaSuccess) return set_error(MANSY_E_CUDA, "cudaMemcpyToSymbol failed");
  return MANSY_OK;
#else
  (void)progress_dev;
  return set_error(MANSY_E_STATE, "library was built without MANSY_MBAR_WATCHDOG");
#endif
}

int mansy_policy_tc_set_split(mansy_policy_t p, int32_t split) {
  if (!p) return set_error(MANSY_E_INVALID, "NULL argument");
  if (!p->tc) return set_error(MANSY_E_STATE, "tensor-core state unavailable");
  if (split != 0 && split != 1 && split != 4) return set_error(MANSY_E_INVALID, "split must be 0 (auto), 1 or 4");
  p->tc->split = split;
  return MANSY_OK;
}

}  // extern "C"
