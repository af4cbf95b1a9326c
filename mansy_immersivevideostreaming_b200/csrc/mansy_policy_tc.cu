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
// write the 128x128 feature tile back to shared memory in the canonical K-major SWIZZLE_128B
// layout, where it becomes the A operand of layer 2: D2[128 x 256] += feat_b * Wfc_b^T (actor.fc
// and critic.fc stacked; the shared FeatureNet is evaluated once).  TMEM: 2 x 128 + 256 = 512
// columns.  Heads (128 -> 15, 128 -> 1), softmax and the categorical sample run in the final
// epilogue from registers with the head weights as constant-bank operands.
//
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM lane quarter = warp & 3).  Both the producer and the issuer walk
// the same host-built job list (c_tc[slot].jobs); shared-memory stages are recycled through
// full/empty mbarriers, accumulators through tcgen05.commit barriers.
#include <cuda.h>

#include <cstring>
#include <mutex>
#include <new>

#include "mansy_policy.cuh"

namespace mansy {

constexpr int kTcSlots = 3;            // policies with live tensor-core state per process
constexpr int kTcMaxJobs = 96;
constexpr int kTcThreads = 192;
constexpr int kTcStages = 3;
constexpr uint32_t kStageBytes = 32768;   // L1 job: A box 16 KB + W1 box 16 KB;  L2 job: Wfc box 32 KB
constexpr uint32_t kBoxBytes = 16384;     // 128 rows x 128 B
constexpr uint32_t kFeatBytes = 65536;    // 128 rows x 128 floats = 4 boxes
constexpr uint32_t kTcSmemBytes = kTcStages * kStageBytes + 2 * kFeatBytes + 256 /*barriers*/ + 1024 /*alignment*/;

enum : uint8_t { kJobL1 = 0, kJobL2 = 1 };
enum : uint8_t { kFlagFirst = 1, kFlagLast = 2, kFlagTileFirstL2 = 4, kFlagTileLastL2 = 8 };

struct TcJob {
  uint8_t type;     // kJobL1 / kJobL2
  uint8_t a_box;    // L1: observation box (32 floats)            L2: -
  uint8_t w_box;    // L1: box of the layer-1 weight image         L2: -
  uint8_t s_lo;     // L1: first K-step inside the box (0..3)      L2: -
  uint8_t s_hi;     // L1: one past the last K-step                L2: -
  uint8_t slot;     // branch in processing order (D1 / feature buffer = slot & 1)
  uint8_t flags;
  uint8_t chunk;    // L2: 32-float chunk of the branch's 128 features (0..3)
};

struct TcConst {
  float wout_t[kHidden][16];   // [j][o]: o < 15 actor.out[o][j], o == 15 critic.out[0][j]
  float bias2[2 * kHidden];    // actor.fc bias | critic.fc bias
  float bias1[kMaxBranches][kHidden];   // processing order
  float bout[16];
  TcJob jobs[kTcMaxJobs];
  int32_t n_jobs, n_branches, residual_slot, softmax;
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
};

struct TcState {
  int slot = -1;
  CUtensorMap map_w1, map_wfc;
  int obs_floats = 0;          // 784 / 400
  int n_branches = 0;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE_%=;\n"
      "bra MBAR_WAIT_%=;\n"
      "MBAR_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 | SBO << 32 | version 1 << 46 | layout 2 << 61).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ULL << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B TF32, both K-major, M = 128.
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
template <int SLOT>
__global__ void __launch_bounds__(kTcThreads, 1)
policy_tc_kernel(const __grid_constant__ CUtensorMap map_obs, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_wfc, const TcArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const TcConst &K = c_tc[SLOT];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage0 = base;
  const uint32_t feat0 = base + kTcStages * kStageBytes;
  const uint32_t bars = feat0 + 2 * kFeatBytes;
  // barrier map (8 bytes each)
  const uint32_t bar_full = bars;                    // [kTcStages]   TMA -> MMA
  const uint32_t bar_empty = bars + 8 * kTcStages;   // [kTcStages]   MMA (commit) -> TMA
  const uint32_t bar_d1_full = bars + 48;            // [2]           MMA (commit) -> epilogue
  const uint32_t bar_feat_full = bars + 64;          // [2]           epilogue (128 arrivals) -> MMA
  const uint32_t bar_feat_empty = bars + 80;         // [2]           MMA (commit) -> epilogue
  const uint32_t bar_d2_full = bars + 96;            //               MMA (commit) -> epilogue
  const uint32_t bar_d2_empty = bars + 104;          //               epilogue (128 arrivals) -> MMA
  const uint32_t tmem_slot = bars + 112;             // uint32 written by tcgen05.alloc

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_d1_full + 8 * b, 1);
      mbar_init(bar_feat_full + 8 * b, 128);
      mbar_init(bar_feat_empty + 8 * b, 1);
    }
    mbar_init(bar_d2_full, 1);
    mbar_init(bar_d2_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_obs) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wfc) : "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns (D1[0] 0..127, D1[1] 128..255, D2 256..511)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int n_jobs = K.n_jobs;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        for (int j = 0; j < n_jobs; ++j, ++it) {
          const TcJob job = K.jobs[j];
          const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t dst = stage0 + s * kStageBytes, full = bar_full + 8 * s;
          mbar_expect_tx(full, kStageBytes);
          if (job.type == kJobL1) {
            tma_load_2d(dst, &map_obs, job.a_box * 32, tile * 128, full);
            tma_load_2d(dst + kBoxBytes, &map_w1, job.w_box * 32, 0, full);
          } else {
            tma_load_2d(dst, &map_wfc, job.slot * kHidden + job.chunk * 32, 0, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t kIdesc128 = idesc_tf32(128), kIdesc256 = idesc_tf32(256);
      uint32_t it = 0, feat_use[2] = {0, 0}, tile_i = 0;
      for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_i) {
        for (int j = 0; j < n_jobs; ++j, ++it) {
          const TcJob job = K.jobs[j];
          const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
          const uint32_t buf = job.slot & 1u;
          const uint32_t st_addr = stage0 + s * kStageBytes;
          if (job.type == kJobL1) {
            // D1[buf] was drained by the epilogue of branch slot-2: its feat_full was awaited before
            // the layer-2 MMAs of that branch, which precede this job in the list.
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t d1 = tmem_base + buf * 128u;
            for (int st = job.s_lo; st < job.s_hi; ++st) {
              const uint64_t ad = smem_desc_sw128(st_addr + st * 32);
              const uint64_t bd = smem_desc_sw128(st_addr + kBoxBytes + st * 32);
              umma_tf32(d1, ad, bd, kIdesc128, ((job.flags & kFlagFirst) && st == job.s_lo) ? 0u : 1u);
            }
            umma_commit(bar_empty + 8 * s);
            if (job.flags & kFlagLast) umma_commit(bar_d1_full + 8 * buf);
          } else {
            if (job.chunk == 0) {   // feature tile of this branch written by the epilogue warps
              mbar_wait(bar_feat_full + 8 * buf, feat_use[buf] & 1u);
              ++feat_use[buf];
            }
            if ((job.flags & kFlagTileFirstL2) && tile_i > 0) mbar_wait(bar_d2_empty, (tile_i - 1) & 1u);
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t d2 = tmem_base + 256u;
            const uint32_t fa = feat0 + buf * kFeatBytes + job.chunk * kBoxBytes;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = smem_desc_sw128(fa + ks * 32);
              const uint64_t bd = smem_desc_sw128(st_addr + ks * 32);
              umma_tf32(d2, ad, bd, kIdesc256, ((job.flags & kFlagTileFirstL2) && ks == 0) ? 0u : 1u);
            }
            umma_commit(bar_empty + 8 * s);
            if (job.flags & kFlagLast) umma_commit(bar_feat_empty + 8 * buf);
            if (job.flags & kFlagTileLastL2) umma_commit(bar_d2_full);
          }
        }
      }
    }
  } else {
    // ===== epilogue warps (2..5) =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // row (environment) inside the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    const int nb = K.n_branches;
    uint32_t d1_use[2] = {0, 0}, feat_use[2] = {0, 0}, tile_i = 0;
    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_i) {
      const int env = tile * 128 + r;
      const bool live = env < A.n;
      for (int i = 0; i < nb; ++i) {
        const uint32_t buf = i & 1u;
        mbar_wait(bar_d1_full + 8 * buf, d1_use[buf] & 1u);
        ++d1_use[buf];
        tc_fence_after();
        if (feat_use[buf] > 0) mbar_wait(bar_feat_empty + 8 * buf, (feat_use[buf] - 1) & 1u);
        ++feat_use[buf];
        const uint32_t fb = feat0 + buf * kFeatBytes + row_off;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          tmem_ld32(tmem_base + lane_addr + buf * 128u + c * 32u, v);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = leaky(v[jj] + K.bias1[i][c * 32 + jj]);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            st_shared_f4(fb + c * kBoxBytes + (((uint32_t)j4 ^ sw) << 4), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2],
                         v[4 * j4 + 3]);
          if (A.feat_dbg && live) {
            float *dst = A.feat_dbg + (size_t)env * (nb * kHidden) + i * kHidden + c * 32;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) dst[jj] = v[jj];
          }
        }
        fence_async_smem();      // generic-proxy writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        mbar_arrive(bar_feat_full + 8 * buf);
      }

      // ---- final epilogue: hidden layer activation, residual, heads, sample -----------------
      mbar_wait(bar_d2_full, tile_i & 1u);
      tc_fence_after();
      float acc[16];
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[o] = K.bout[o];
      const int rs = K.residual_slot;
      const uint32_t rb = feat0 + (uint32_t)(rs & 1) * kFeatBytes + row_off;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float va[32], vc[32], rr[32];
        tmem_ld32(tmem_base + lane_addr + 256u + c * 32u, va);
        tmem_ld32(tmem_base + lane_addr + 384u + c * 32u, vc);
        if (c == 3) {            // D2 fully read: the next tile's layer 2 may overwrite it
          tc_fence_before();
          mbar_arrive(bar_d2_empty);
        }
        if (rs >= 0) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = ld_shared_f4(rb + c * kBoxBytes + (((uint32_t)j4 ^ sw) << 4));
            rr[4 * j4] = t.x; rr[4 * j4 + 1] = t.y; rr[4 * j4 + 2] = t.z; rr[4 * j4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) rr[jj] = 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int j = c * 32 + jj;
          va[jj] = leaky(va[jj] + K.bias2[j]) + rr[jj];                 // mansy.py:65: fc(features) + qoe_features
          vc[jj] = leaky(vc[jj] + K.bias2[kHidden + j]) + rr[jj];       // mansy.py:79
#pragma unroll
          for (int o = 0; o < kActions; ++o) acc[o] = fmaf(va[jj], K.wout_t[j][o], acc[o]);
          acc[15] = fmaf(vc[jj], K.wout_t[j][15], acc[15]);
        }
        if (A.hid_dbg && live) {
          float *dst = A.hid_dbg + (size_t)env * 256 + c * 32;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) { dst[jj] = va[jj]; dst[kHidden + jj] = vc[jj]; }
        }
      }
      if (live) {
        float p[kActions];
#pragma unroll
        for (int o = 0; o < kActions; ++o) p[o] = acc[o];
        if (K.softmax) {          // simple_rl.py:48: the actor returns probabilities
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
          categorical_sample(p, K.softmax, A.seed, (uint64_t)(A.env_offset + env), (uint64_t)A.step, act, lp);
          A.actions[env] = act;
          if (A.logp) A.logp[env] = lp;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
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

int tc_create(mansy_policy *p, const mansy_policy_weights_t *w) {
  const bool is_mansy = w->kind == MANSY_OBS_MANSY;
  const int nb = is_mansy ? 10 : 5;
  const int obs_floats = is_mansy ? MANSY_OBS_MANSY_STRIDE : MANSY_OBS_SIMPLE_STRIDE;
  // processing order: big branches first, the residual branch last (its feature tile must still
  // be in shared memory when the heads run); `extra` marks the branch whose K-step is shared.
  static const BranchPlan mansy_plan[10] = {
      {1, 8, 320, false},  {2, 328, 320, false}, {3, 648, 64, false}, {0, 0, 8, false},   {4, 728, 8, false},
      {5, 736, 8, false},  {6, 744, 8, false},   {7, 752, 8, false},  {8, 779, 1, true},  {9, 776, 3, false}};
  static const BranchPlan simple_plan[5] = {
      {1, 8, 320, false}, {4, 328, 64, false}, {0, 0, 8, false}, {2, 394, 1, true}, {3, 392, 2, false}};
  const BranchPlan *plan = is_mansy ? mansy_plan : simple_plan;
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
        wfcimg[(size_t)(kHidden + j) * F + i * kHidden + f] = w->critic_fc_w[(size_t)j * F + b.canon * kHidden + f];
      }
    }
  }
  for (int j = 0; j < kHidden; ++j) {
    for (int o = 0; o < kActions; ++o) hc->wout_t[j][o] = w->actor_out_w[(size_t)o * kHidden + j];
    hc->wout_t[j][15] = w->critic_out_w[j];
    hc->bias2[j] = w->actor_fc_b[j];
    hc->bias2[kHidden + j] = w->critic_fc_b[j];
  }
  for (int o = 0; o < kActions; ++o) hc->bout[o] = w->actor_out_b[o];
  hc->bout[15] = w->critic_out_b[0];
  hc->n_branches = nb;
  hc->softmax = is_mansy ? 0 : 1;
  hc->residual_slot = is_mansy ? nb - 1 : -1;

  // job list: L1(p0), L1(p1), L2(p0), L1(p2), L2(p1), ..., L1(p_last), L2(p_last-1), L2(p_last)
  int nj = 0;
  auto push_l1 = [&](int i) {
    const BranchPlan &b = plan[i];
    const int lo = b.off / 8, hi = (b.off + b.k + 7) / 8;      // K-steps of 8 floats
    for (int box = lo / 4; box <= (hi - 1) / 4; ++box) {
      TcJob &j = hc->jobs[nj++];
      j.type = kJobL1; j.a_box = (uint8_t)box; j.w_box = (uint8_t)(b.extra ? main_boxes : box);
      j.s_lo = (uint8_t)((lo > box * 4 ? lo : box * 4) - box * 4);
      j.s_hi = (uint8_t)((hi < box * 4 + 4 ? hi : box * 4 + 4) - box * 4);
      j.slot = (uint8_t)i; j.chunk = 0;
      j.flags = (uint8_t)((box == lo / 4 ? kFlagFirst : 0) | (box == (hi - 1) / 4 ? kFlagLast : 0));
    }
  };
  auto push_l2 = [&](int i) {
    for (int c = 0; c < 4; ++c) {
      TcJob &j = hc->jobs[nj++];
      j.type = kJobL2; j.a_box = j.w_box = j.s_lo = j.s_hi = 0; j.slot = (uint8_t)i; j.chunk = (uint8_t)c;
      j.flags = (uint8_t)((c == 3 ? kFlagLast : 0) | (i == 0 && c == 0 ? kFlagTileFirstL2 : 0) |
                          (i == nb - 1 && c == 3 ? kFlagTileLastL2 : 0));
    }
  };
  push_l1(0);
  for (int i = 1; i < nb; ++i) { push_l1(i); push_l2(i - 1); }
  push_l2(nb - 1);
  hc->n_jobs = nj;

  int rc = MANSY_OK;
  float *d_w1 = nullptr, *d_wfc = nullptr;
  if (cudaMalloc(&d_w1, w1img.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_wfc, wfcimg.size() * sizeof(float)) != cudaSuccess)
    rc = set_error(MANSY_E_NOMEM, "cudaMalloc failed (tensor-core weight images)");
  if (d_w1) p->allocs.push_back(d_w1);
  if (d_wfc) p->allocs.push_back(d_wfc);
  if (!rc && (cudaMemcpy(d_w1, w1img.data(), w1img.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpy(d_wfc, wfcimg.data(), wfcimg.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpyToSymbol(c_tc, hc, sizeof(TcConst), (size_t)st->slot * sizeof(TcConst)) != cudaSuccess))
    rc = set_error(MANSY_E_CUDA, "upload of the tensor-core weight images failed");
  delete hc;
  if (!rc) rc = make_map(&st->map_w1, d_w1, (uint64_t)w1_cols, kHidden, (uint64_t)w1_cols, 128);
  if (!rc) rc = make_map(&st->map_wfc, d_wfc, (uint64_t)F, 256, (uint64_t)F, 256);
  if (rc) {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    g_slot_used[st->slot] = false;
    delete st;
    return rc;
  }
  p->tc = st;
  return MANSY_OK;
}

void tc_destroy(mansy_policy *p) {
  if (!p || !p->tc) return;
  {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    g_slot_used[p->tc->slot] = false;
  }
  delete p->tc;
  p->tc = nullptr;
}

}  // namespace mansy

using namespace mansy;

extern "C" {

int mansy_policy_forward_tc(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                            float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                            int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, void *stream) {
  if (!p || !obs_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (!p->tc) return set_error(MANSY_E_STATE, std::string("tensor-core state unavailable: ") + mansy_last_error());
  if (n < 0) return set_error(MANSY_E_INVALID, "n < 0");
  if (obs_stride < p->tc->obs_floats || (obs_stride & 3))
    return set_error(MANSY_E_INVALID, "obs_stride must be >= the padded row length and a multiple of 4 floats");
  if (reinterpret_cast<uintptr_t>(obs_dev) & 15) return set_error(MANSY_E_INVALID, "obs must be 16-byte aligned");
  if (logits_dev && (reinterpret_cast<uintptr_t>(logits_dev) & 15)) return set_error(MANSY_E_INVALID, "logits must be 16-byte aligned");
  if (n == 0) return MANSY_OK;
  CUtensorMap map_obs;
  int rc = make_map(&map_obs, obs_dev, (uint64_t)p->tc->obs_floats, (uint64_t)n, (uint64_t)obs_stride, 128);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.n_tiles = (n + 127) / 128;
  a.logits = logits_dev; a.value = value_dev; a.actions = actions_dev; a.logp = logp_dev;
  a.feat_dbg = feat_dbg_dev; a.hid_dbg = hid_dbg_dev;
  a.seed = seed; a.step = step; a.env_offset = env_offset;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm < 1) n_sm = 148;
  }
  const int grid = a.n_tiles < n_sm ? a.n_tiles : n_sm;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
#define MANSY_TC_LAUNCH(SLOT)                                                                                       \
  do {                                                                                                              \
    static bool attr_done = false;                                                                                  \
    if (!attr_done) {                                                                                               \
      e = cudaFuncSetAttribute(policy_tc_kernel<SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes); \
      attr_done = (e == cudaSuccess);                                                                               \
    }                                                                                                               \
    if (e == cudaSuccess)                                                                                           \
      policy_tc_kernel<SLOT><<<grid, kTcThreads, kTcSmemBytes, s>>>(map_obs, p->tc->map_w1, p->tc->map_wfc, a);      \
  } while (0)
  switch (p->tc->slot) {
    case 0: MANSY_TC_LAUNCH(0); break;
    case 1: MANSY_TC_LAUNCH(1); break;
    default: MANSY_TC_LAUNCH(2); break;
  }
#undef MANSY_TC_LAUNCH
  if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("policy_tc_kernel attribute: ") + cudaGetErrorString(e));
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MANSY_E_CUDA, std::string("policy_tc_kernel launch: ") + cudaGetErrorString(e));
  return MANSY_OK;
}

}  // extern "C"
