// mansy_mtio.cu -- MTIO viewport-prediction transformer, inference path, for sm_100a.
//
// Reference: ViewportTransformerMTIO.sample viewport_prediction/models/mtio.py:106-133 (called by predict.py:27),
// _process_src_tgt mtio.py:135-148, ViewportEmbedding / PositionalEncoding mtio.py:11-45, Transformer.forward +
// DistillLayer viewport_prediction/models/customized_transformer.py:13-36,52-70, torch.nn.Transformer
// (post-norm, ReLU, 8 heads: the constructor defaults the reference passes), to_position_normalized_cartesian
// viewport_prediction/utils/common.py:61-70.  The reference runs these layers in TF32
// (torch.set_float32_matmul_precision('high'), predict.py:100).
//
// Structure.  In eval mode the encoder output does not depend on the decoding step, so it is computed once per
// sample; the decoder is causal, so step t only computes token t and keeps the keys / values of tokens 0..t-1
// per layer in HBM (the reference re-runs encoder and decoder on the whole prefix every step, mtio.py:120-123:
// ~10x the FLOPs for the same function).  What is left is a chain of [rows x 512] x [512 x N] products shared
// by all samples:
//   * mtio_gemm_kernel<BN, MT, EPI>: tcgen05 kind::tf32.  A rows and torch-layout weight rows ([out][in] = K-major) are
//     TMA-loaded as [128 x 32-float] SWIZZLE_128B boxes into a 2-3-stage ring (one producer warp per stage half), one
//     elected lane issues M128 x N256 x K8 MMAs into TMEM, four epilogue warps (one per TMEM lane quarter, thread =
//     output row) post-process and stage [128 x 32] output boxes in swizzled shared memory for TMA stores.  Tiles:
//        <256, 2> [256 rows x 256 columns]: two M128 accumulators share every weight box -- all GEMMs without
//                 LayerNorm (EPI_NONE / EPI_RELU / EPI_ELU: + bias (+ activation));
//        <256, 1> [128 x 256], two CTAs per SM: small batches, and EPI_LN2, the LayerNorm GEMM as a 2-CTA cluster whose
//                 CTAs own the two column halves of the same rows: + bias + residual (TMA-prefetched boxes), per-row
//                 (sum, centred squares) swapped through distributed shared memory, normalise -- the post-norm
//                 "x = norm(x + sublayer(x))" of every transformer sub-block in the producing kernel;
//        <512, 1> [128 x 512] EPI_LN: the same epilogue with the whole d_model row in one CTA (MANSY_MTIO_LN_SPLIT=0).
//     Output columns are routed by 512-column segment (tensor map per segment): the fused QKV projection writes q to a
//     scratch row and k / v straight into step t of the per-layer cache.  Passes of >= 1024 samples run as two lanes
//     (two streams, two launching threads).  DESIGN.md 4.7 has the measurements behind these choices.
//   * mtio_attn_kernel: one warp per (sample, head); <= 32 keys (5 source tokens, 3 distilled memory tokens,
//     <= 31 cached target tokens): lane = key for the scores, lane = 2 of the 64 head dims for the output.
//   * mtio_head_kernel: final decoder LayerNorm + predictor + sigmoid + MTIO-head ensemble + wrap into the unit
//     square + embedding / positional encoding of the NEXT token, one warp per sample, one launch per step.
//   * small CUDA-core kernels for the 6-wide embedding, the final encoder norm, the circular-conv im2col of the
//     DistillLayer (its BatchNorm is folded into the conv weights at create; the conv is a K = 1536 GEMM with
//     an ELU epilogue) and its max-pool.
// MANSY_MTIO_FP32 swaps the GEMMs for an exact-fp32 CUDA-core kernel (parity anchor, like mansy_policy.cu).
#include <cuda.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "mansy_b200.h"
#include "mansy_policy.cuh"
#include "mansy_tc.cuh"

namespace mansy {

constexpr int kD = 512;        // d_model = dim_feedforward (predict.py:73-74)
constexpr int kMtioHeads = 8;  // nn.Transformer default nhead (customized_transformer.py:40)
constexpr int kDh = kD / kMtioHeads;
constexpr int kTok = 6;        // in_channel 2 x 3 MTIO heads (mtio.py:49,55)
constexpr float kLnEps = 1e-5f;

enum { EPI_NONE = 0, EPI_RELU = 1, EPI_ELU = 2, EPI_LN = 3, EPI_LN2 = 4 };   // LN2: LayerNorm row split over a 2-CTA cluster

struct GemmArgs {
  int32_t M, N, K;
  const float *bias;       // [N]
  float *out[3];           // output segment s covers columns [512 s, 512 s + 512)
  int64_t out_ld[3];       // floats between rows of the segment
  const float *res;        // EPI_LN: residual rows [M][res_ld]
  int64_t res_ld;
  const float *gamma, *beta;
  int32_t w_box_rows;      // rows of the weight tensor map's box: 128, or 256 (two adjacent boxes per TMA operation)
  long long *timeline;     // profiling hook (MANSY_MTIO_TIMELINE): SM-clock stamps of CTA (0, 0), see mtio_timeline_dump
};

// ------------------------------------------------------------------------------------------
// tcgen05 GEMM
// ------------------------------------------------------------------------------------------
constexpr int kGemmMmaWarp = 4;     // warps 0..3: epilogue (TMEM lane quarter = warp), 4: MMA issuer, 5..: TMA producers, last: residual producer
constexpr int kGemmProd0 = 5;
constexpr uint32_t kABoxBytes = 16384;     // 128 rows x 128 B
constexpr int kResRing = 3;                // residual boxes prefetched while the main loop runs (BN = 512 only)
constexpr uint32_t kGemmBarBytes = 256;

// Tile shapes.  What bounds these products is the operand stream out of L2 (ncu: tensor pipe ~32 % busy with
// [128 x 256] tiles while L2 -> SM traffic sits at ~7.6 TB/s), so the tile is made as large as tensor memory allows:
//   <256, 2>: [256 rows x 256 columns] = two M128 accumulators sharing every weight box (64 KB per K-chunk of 32
//             for 4.2 MFLOP = 65 FLOP/B), 3 stages, one CTA per SM -- every GEMM without LayerNorm;
//   <256, 1>: [128 x 256], 2 stages, two CTAs per SM -- batches of <= 128 rows;
//   <512, 1>: [128 x 512], the LayerNorm epilogue's full-row tile (52 FLOP/B).
template <int BN, int MT>
struct GemmCfg {
  static constexpr int kStages = MT == 2 ? 3 : 2;
  static constexpr uint32_t kStage = MT * kABoxBytes + BN * 128u;
  static constexpr int kResWarp = kGemmProd0 + 2 * kStages;
  static constexpr int kThreads = 32 * (kResWarp + 1);
  static constexpr uint32_t kRing = BN == 512 ? kResRing * kABoxBytes : 0u;
  static constexpr uint32_t kSmem = 1024u + kStages * kStage + kRing + kGemmBarBytes + 3u * BN * 4u + 1024u /* LN2 statistics */;
  static constexpr int kCtasPerSm = (BN == 256 && MT == 1) ? 2 : 1;
  static constexpr int kOutSlots = BN == 512 ? 8 : 4;       // output staging boxes carved out of the operand stages
};

struct OutMaps { CUtensorMap m[3]; };      // one tensor map per 512-column output segment

// TMA store of a [128 x 32-float] SWIZZLE_128B box from shared memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Cluster launch (runtime attribute): the CTAs of a cluster own consecutive M tiles of the same N tile, so they need the
// same weight boxes.  Box b of a stage is fetched by CTA rank b % cluster_size and multicast into every CTA of the
// cluster (same shared-memory offset, each CTA's own `full` barrier gets the bytes); a stage is free again when the MMAs
// of ALL CTAs have read it (commit multicast onto every CTA's `empty` barrier, which counts cluster_size arrivals).
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

template <int BN, int MT, int EPI>
__global__ void __launch_bounds__(GemmCfg<BN, MT>::kThreads, GemmCfg<BN, MT>::kCtasPerSm)
mtio_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_res, const __grid_constant__ OutMaps maps_out, const GemmArgs g) {
  using Cfg = GemmCfg<BN, MT>;
  static_assert((BN == 256 && (MT == 1 || MT == 2)) || (BN == 512 && MT == 1), "tile shape");
  static_assert(EPI != EPI_LN || BN == 512, "the LayerNorm epilogue needs the whole d_model row in one tile");
  static_assert(EPI != EPI_LN2 || (BN == 256 && MT == 1), "split LayerNorm: two [128 x 256] halves in a 2-CTA cluster");
  constexpr bool kLn = EPI == EPI_LN || EPI == EPI_LN2;        // epilogues with residual + LayerNorm
  constexpr bool kLn2 = EPI == EPI_LN2;
  constexpr int kBoxes = BN / 32;                                // residual / output boxes of 32 columns per tile row block
  constexpr int kRingBoxes = (int)(Cfg::kRing / kABoxBytes);     // residual boxes with a buffer of their own (BN = 512: 3)
  constexpr int kStageBoxes = (int)(Cfg::kStages * Cfg::kStage / kABoxBytes);   // residual boxes that fit the freed operand stages
  constexpr int kReuse = kBoxes - kRingBoxes - kStageBoxes > 0 ? kBoxes - kRingBoxes - kStageBoxes : 0;   // boxes that wait for a consumed buffer
  static_assert(BN * MT <= 512, "tensor memory columns");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kStage = Cfg::kStage;
  constexpr int kStages = Cfg::kStages;
  constexpr int NB = BN / 128;                   // weight boxes of 128 rows per stage
  constexpr uint32_t kWOff = MT * kABoxBytes;    // weight boxes follow the A boxes inside a stage
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t ring = base + kStages * kStage;                  // LayerNorm epilogue: residual boxes 0..2 (and 13..15)
  const uint32_t bars = ring + Cfg::kRing;
  const uint32_t bar_full = bars;                // [stage][half], <= 3 stages
  const uint32_t bar_empty = bars + 48;          // [stage]
  const uint32_t bar_d = bars + 72;
  const uint32_t tmem_slot = bars + 80;
  const uint32_t bar_res_full = bars + 88;       // [16] residual box c landed (single use)
  const uint32_t bar_ring_free = bars + 216;     // [kResRing] ring box consumed by the 128 epilogue threads
  float *s_bias = reinterpret_cast<float *>(smem_raw + (bars + kGemmBarBytes - raw));
  float *s_gamma = s_bias + BN, *s_beta = s_gamma + BN;
  const uint32_t xch = bars + kGemmBarBytes + 3u * BN * 4u;      // LN2: [128 rows] x (sum, centred squares) written by the peer CTA

  const int warp = threadIdx.x >> 5;
  const uint32_t crank = cluster_rank(), csize_hw = cluster_size();
  // LN2: the two CTAs of a cluster own the two 256-column halves of the same 128 rows (nothing is multicast)
  const uint32_t csize = kLn2 ? 1u : csize_hw;
  const int m0 = kLn2 ? (int)(blockIdx.x >> 1) * 128 : blockIdx.x * (128 * MT);
  const int n0 = kLn2 ? (int)crank * BN : blockIdx.y * BN;
  const int nk = g.K >> 5;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 16 * s, 1);
      mbar_init(bar_full + 16 * s + 8, 1);
      mbar_init(bar_empty + 8 * s, csize);
    }
    mbar_init(bar_d, 1);
    if (kLn) {
      for (int c = 0; c < kBoxes; ++c) mbar_init(bar_res_full + 8 * c, 1);
      for (int c = 0; c < kReuse; ++c) mbar_init(bar_ring_free + 8 * c, 128);
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_res) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps_out.m[n0 >> 9]) : "memory");
  }
  if (warp == kGemmMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(BN * MT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < 4) {
    for (int i = threadIdx.x; i < BN; i += 128) {
      s_bias[i] = g.bias[n0 + i];
      if (kLn) { s_gamma[i] = g.gamma[n0 + i]; s_beta[i] = g.beta[n0 + i]; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_barrier();       // every CTA's barriers exist before a peer's multicast / commit can reach them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const bool tl = g.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  if (tl && threadIdx.x == 0) g.timeline[190] = clock64();        // prologue done

  if (warp == Cfg::kResWarp) {
    // ===== residual producer (LayerNorm epilogue): the [128 x 512] residual tile as 16 boxes of 32 columns.  Boxes 0..2
    // land in the ring while the main loop runs, 3..12 in the operand stages once the last MMA has retired, 13..15 in
    // the ring again as the epilogue frees it. =====
    if (kLn) {
      auto buffer_of = [&](int c) -> uint32_t {      // residual box c: own ring box, a freed operand-stage box, or a consumed buffer
        const int j = c < kRingBoxes + kStageBoxes ? c : c - kRingBoxes - kStageBoxes;
        return j < kRingBoxes ? ring + j * kABoxBytes : base + (j - kRingBoxes) * kABoxBytes;
      };
      if (elect_one()) {
        for (int c = 0; c < kRingBoxes; ++c) {
          mbar_expect_tx(bar_res_full + 8 * c, kABoxBytes);
          tma_load_2d(buffer_of(c), &map_res, n0 + c * 32, m0, bar_res_full + 8 * c);
        }
      }
      __syncwarp();
      mbar_wait(bar_d, 0);
      if (elect_one()) {
        for (int c = kRingBoxes; c < kBoxes - kReuse; ++c) {
          mbar_expect_tx(bar_res_full + 8 * c, kABoxBytes);
          tma_load_2d(buffer_of(c), &map_res, n0 + c * 32, m0, bar_res_full + 8 * c);
        }
      }
      __syncwarp();
      for (int c = kBoxes - kReuse; c < kBoxes; ++c) {
        mbar_wait(bar_ring_free + 8 * (c - (kBoxes - kReuse)), 0);
        if (elect_one()) {
          mbar_expect_tx(bar_res_full + 8 * c, kABoxBytes);
          tma_load_2d(buffer_of(c), &map_res, n0 + c * 32, m0, bar_res_full + 8 * c);
        }
        __syncwarp();
      }
    }
    if (kLn2) cluster_barrier();           // every thread of both CTAs meets here once (statistics exchange below)
  } else if (warp >= kGemmProd0) {
    // ===== TMA producers: warp p fills half (p & 1) of stage (p >> 1).
    // MT = 1: half 0 = A box + first NB/2 weight boxes, half 1 = the other weight boxes;  MT = 2: half 0 = both A boxes,
    // half 1 = both weight boxes (32 KB each). =====
    const int p = warp - kGemmProd0, s = p >> 1, h = p & 1;
    const uint32_t dst = base + s * kStage, full = bar_full + 16 * s + 8 * h;
    for (int it = s; it < nk; it += kStages) {
      mbar_wait(bar_empty + 8 * s, ((uint32_t)(it / kStages) & 1u) ^ 1u);
      if (tl && h == 0 && (threadIdx.x & 31) == 0 && it < 60) g.timeline[it] = clock64();           // stage free: TMA issue
      if (elect_one()) {
        auto load_w = [&](int b) {          // weight box b (rows n0 + 128 b ..) of this K-chunk
          const uint32_t at = dst + kWOff + kABoxBytes * b;
          if (g.w_box_rows == 256) {        // the tensor map's box is [256 x 32 floats]: boxes (b, b + 1) in one operation
            if (!(b & 1)) tma_load_2d(at, &map_w, it * 32, n0 + b * 128, full);
          } else if (csize == 1) {
            tma_load_2d(at, &map_w, it * 32, n0 + b * 128, full);
          } else if ((uint32_t)b % csize == crank) {
            tma_load_2d_mc(at, &map_w, it * 32, n0 + b * 128, full, cmask);
          }
        };
        if (MT == 2) {
          mbar_expect_tx(full, 2 * kABoxBytes);
          if (h == 0) {
            tma_load_2d(dst, &map_a, it * 32, m0, full);
            tma_load_2d(dst + kABoxBytes, &map_a, it * 32, m0 + 128, full);
          } else {
            load_w(0);
            load_w(1);
          }
        } else if (NB == 2 && g.w_box_rows == 256) {      // one 256-row operation brings both weight boxes: all on half 0
          if (h == 0) {
            mbar_expect_tx(full, 3 * kABoxBytes);
            tma_load_2d(dst, &map_a, it * 32, m0, full);
            load_w(0);
          } else {
            mbar_arrive(full);
          }
        } else if (h == 0) {
          mbar_expect_tx(full, kABoxBytes * (1 + NB / 2));
          tma_load_2d(dst, &map_a, it * 32, m0, full);
#pragma unroll
          for (int b = 0; b < NB / 2; ++b) load_w(b);
        } else {
          mbar_expect_tx(full, kABoxBytes * (NB / 2));
#pragma unroll
          for (int b = NB / 2; b < NB; ++b) load_w(b);
        }
      }
      __syncwarp();
    }
    if (kLn2) cluster_barrier();
  } else if (warp == kGemmMmaWarp) {
    // ===== MMA issuer: per stage four K8 steps of M128 x N256 MMAs =====
    constexpr uint32_t kIdesc = idesc_tf32(256);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nk; ++it) {
      const uint32_t a_lo = smem_desc_lo(base + s * kStage), b_lo = smem_desc_lo(base + s * kStage + kWOff);
      if (BN == 256) {
        mbar_wait(bar_full + 16 * s, ph);
        mbar_wait(bar_full + 16 * s + 8, ph);
        if (tl && (threadIdx.x & 31) == 0 && it < 60) g.timeline[60 + it] = clock64();              // operands landed
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32(tmem_base + 256u * mt, make_desc(a_lo + mt * (kABoxBytes >> 4) + 2 * ks), make_desc(b_lo + 2 * ks), kIdesc,
                        (it > 0 || ks > 0) ? 1u : 0u);
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {     // columns 0..255 need half 0 only; 256..511 need the second weight half
          mbar_wait(bar_full + 16 * s + 8 * h, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32(tmem_base + 256u * h, make_desc(a_lo + 2 * ks), make_desc(b_lo + h * ((2 * kABoxBytes) >> 4) + 2 * ks),
                        kIdesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
          __syncwarp();
        }
      }
      if (elect_one()) {
        if (csize == 1) umma_commit(bar_empty + 8 * s);
        else umma_commit_mc(bar_empty + 8 * s, cmask);
        if (it == nk - 1) umma_commit(bar_d);
      }
      __syncwarp();
      if (tl && (threadIdx.x & 31) == 0 && it < 60) g.timeline[120 + it] = clock64();                 // MMAs issued + committed
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    if (kLn2) cluster_barrier();
  } else {
    // ===== epilogue warps 0..3: thread = output row.  Output (and residual) boxes are [128 rows x 128 B] SWIZZLE_128B
    // tiles: row r keeps its 16-byte piece j at r * 128 + ((j ^ (r & 7)) << 4), so the 8 lanes of a shared-memory phase
    // touch 8 different bank groups.  Boxes leave through TMA stores (rows >= M are clipped by the tensor map). =====
    const int r = warp * 32 + (threadIdx.x & 31);
    const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t row_off = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
    const CUtensorMap *omap = &maps_out.m[n0 >> 9];
    const int col0 = n0 & 511;
    constexpr int NS = Cfg::kOutSlots;
    mbar_wait(bar_d, 0);
    tc_fence_after();
    if (tl && threadIdx.x == 0) g.timeline[191] = clock64();        // accumulators complete
    if (!kLn) {
#pragma unroll 1
      for (int ci = 0; ci < MT * (BN / 32); ++ci) {
        const int mt = ci / (BN / 32), c = ci % (BN / 32);
        const uint32_t slot = base + (uint32_t)(ci % NS) * kABoxBytes;
        if (ci >= NS) {
          if (threadIdx.x == 0) tma_store_wait_read<NS - 1>();      // the store issued NS boxes ago has read its slot
          epi_bar_sync();
        }
        float v[32];
        tmem_ld32(ta + mt * 256 + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = v[j] + s_bias[c * 32 + j];
          if (EPI == EPI_RELU) x = fmaxf(x, 0.f);
          if (EPI == EPI_ELU) x = x > 0.f ? x : expm1f(x);
          v[j] = x;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(slot + row_off + ((((uint32_t)j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_async_smem();
        epi_bar_sync();
        if (threadIdx.x == 0) {
          tma_store_2d(omap, slot, col0 + c * 32, m0 + mt * 128);
          tma_store_commit();
        }
      }
      if (threadIdx.x == 0) tma_store_wait_all();
    } else {
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < kBoxes; ++c) {     // pass 1: x = acc + bias + residual, kept in TMEM
        const int j0 = c < kRingBoxes + kStageBoxes ? c : c - kRingBoxes - kStageBoxes;
        const uint32_t buf = j0 < kRingBoxes ? ring + j0 * kABoxBytes : base + (j0 - kRingBoxes) * kABoxBytes;
        float v[32];
        tmem_ld32(ta + c * 32, v);
        mbar_wait(bar_res_full + 8 * c, 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 q = lds128(buf + row_off + ((((uint32_t)j) ^ sw) << 4));
          v[4 * j] += s_bias[c * 32 + 4 * j] + q.x;
          v[4 * j + 1] += s_bias[c * 32 + 4 * j + 1] + q.y;
          v[4 * j + 2] += s_bias[c * 32 + 4 * j + 2] + q.z;
          v[4 * j + 3] += s_bias[c * 32 + 4 * j + 3] + q.w;
        }
        if (c < kReuse) {
          fence_async_smem();            // the buffer is refilled through the async proxy (TMA) once all 128 threads have read it
          mbar_arrive(bar_ring_free + 8 * c);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += v[j];
        tmem_st32(ta + c * 32, v);
      }
      tmem_st_wait();
      float mean = sum * (1.0f / BN);
      float ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < kBoxes; ++c) {     // pass 2: centred sum of squares
        float v[32];
        tmem_ld32(ta + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; ss += d * d; }
      }
      float var = ss * (1.0f / BN);
      if (kLn2) {
        // the row's other 256 columns live in the peer CTA: swap (sum, centred squares) through distributed shared memory
        // and merge the two halves (Chan et al.): M2 = M2_a + M2_b + n_a (mean_a - mean)^2 + n_b (mean_b - mean)^2
        uint32_t peer;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(xch + (uint32_t)r * 8u), "r"(crank ^ 1u));
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(peer), "f"(sum), "f"(ss) : "memory");
        cluster_barrier();
        float osum, oss;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(osum), "=f"(oss) : "r"(xch + (uint32_t)r * 8u) : "memory");
        const float omean = osum * (1.0f / BN);
        const float tmean = (sum + osum) * (0.5f / BN);
        const float da = mean - tmean, db = omean - tmean;
        var = (ss + oss + (float)BN * (da * da + db * db)) * (0.5f / BN);
        mean = tmean;
      }
      const float rstd = 1.0f / sqrtf(var + kLnEps);
      epi_bar_sync();                    // every warp is done reading residual boxes: the stages become output staging
#pragma unroll 1
      for (int c = 0; c < kBoxes; ++c) {     // pass 3: normalise, affine, stage the box, TMA store
        const uint32_t slot = base + (uint32_t)(c % NS) * kABoxBytes;
        if (c >= NS) {
          if (threadIdx.x == 0) tma_store_wait_read<NS - 1>();
          epi_bar_sync();
        }
        float v[32];
        tmem_ld32(ta + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * s_gamma[c * 32 + j] + s_beta[c * 32 + j];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(slot + row_off + ((((uint32_t)j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_async_smem();
        epi_bar_sync();
        if (threadIdx.x == 0) {
          tma_store_2d(omap, slot, col0 + c * 32, m0);
          tma_store_commit();
        }
      }
      if (threadIdx.x == 0) tma_store_wait_all();
    }
  }

  if (tl && threadIdx.x == 0) g.timeline[192] = clock64();          // epilogue done
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_barrier();       // no CTA leaves while a peer's commit may still target its barriers
  if (warp == kGemmMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN * MT) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// exact-fp32 CUDA-core GEMM (parity anchor): C = epi(A W^T + bias [+ residual]); LayerNorm follows in mtio_ln_kernel
// ------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) mtio_sgemm_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ W,
                                                         const GemmArgs g) {
  __shared__ float sa[16][64 + 4], sw[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < g.K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      sa[k][r] = (m0 + r < g.M) ? A[(size_t)(m0 + r) * (size_t)lda + k0 + k] : 0.f;
      sw[k][r] = W[(size_t)(n0 + r) * (size_t)g.K + k0 + k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sa[k][ty * 4 + i]; b[i] = sw[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      float x = acc[i][j] + g.bias[n];
      if (EPI == EPI_RELU) x = fmaxf(x, 0.f);
      if (EPI == EPI_ELU) x = x > 0.f ? x : expm1f(x);
      if (EPI == EPI_LN) x += g.res[(size_t)m * (size_t)g.res_ld + n];
      g.out[n >> 9][(size_t)m * (size_t)g.out_ld[n >> 9] + (n & 511)] = x;
    }
  }
}

// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// LayerNorm of one 512-float row held as 4 float4 per lane (columns i * 128 + lane * 4 .. + 3).
__device__ __forceinline__ void warp_layer_norm(float4 (&x)[4], const float *__restrict__ w, const float *__restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
  const float mean = warp_sum(s) * (1.0f / kD);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = x[i].x - mean, c = x[i].y - mean, d = x[i].z - mean, e = x[i].w - mean;
    ss += (a * a + c * c) + (d * d + e * e);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) * (1.0f / kD) + kLnEps);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 ww = *reinterpret_cast<const float4 *>(w + i * 128 + lane * 4);
    const float4 bb = *reinterpret_cast<const float4 *>(b + i * 128 + lane * 4);
    x[i].x = (x[i].x - mean) * rstd * ww.x + bb.x;
    x[i].y = (x[i].y - mean) * rstd * ww.y + bb.y;
    x[i].z = (x[i].z - mean) * rstd * ww.z + bb.z;
    x[i].w = (x[i].w - mean) * rstd * ww.w + bb.w;
  }
}

// rows of `in` -> LayerNorm -> rows of `out` (may alias), one warp per row
__global__ void __launch_bounds__(256) mtio_ln_kernel(const float *in, float *out, const float *__restrict__ w,
                                                      const float *__restrict__ b, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4 *>(in + row * kD + i * 128 + lane * 4);
  warp_layer_norm(x, w, b, lane);
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4 *>(out + row * kD + i * 128 + lane * 4) = x[i];
}

// ViewportEmbedding + PositionalEncoding (mtio.py:29-45, 11-27): out[row] = W tok + b + pe[pos_base + row % pos_mod].
// in_dim 2: the viewport centre repeated for the 3 MTIO heads (mtio.py:113-116); in_dim 6: a predicted token.
__global__ void __launch_bounds__(128) mtio_embed_kernel(const float *__restrict__ tok, int64_t tok_stride, int in_dim, int64_t rows,
                                                         int pos_mod, int pos_base, const float *__restrict__ emb_w,
                                                         const float *__restrict__ emb_b, const float *__restrict__ pe,
                                                         float *__restrict__ out, float *__restrict__ tok_out, int64_t tok_out_stride) {
  const int64_t row = blockIdx.x;          // one block of 128 threads per row
  const int c4 = (int)threadIdx.x;
  if (row >= rows) return;
  float t[kTok];
  const float *tr = tok + row * tok_stride;
  if (in_dim == 2) {
    const float x = tr[0], y = tr[1];
    t[0] = x; t[1] = y; t[2] = x; t[3] = y; t[4] = x; t[5] = y;
  } else {
#pragma unroll
    for (int o = 0; o < kTok; ++o) t[o] = tr[o];
  }
  if (tok_out && c4 == 0) {
#pragma unroll
    for (int o = 0; o < kTok; ++o) tok_out[row * tok_out_stride + o] = t[o];
  }
  const int pos = pos_base + (int)((uint32_t)blockIdx.x % (uint32_t)pos_mod);
  float4 r = *reinterpret_cast<const float4 *>(emb_b + c4 * 4);
#pragma unroll
  for (int o = 0; o < kTok; ++o) {      // emb_w: transposed [6][512] copy, one 16-byte load per input column
    const float4 w = *reinterpret_cast<const float4 *>(emb_w + o * kD + c4 * 4);
    r.x = fmaf(w.x, t[o], r.x); r.y = fmaf(w.y, t[o], r.y); r.z = fmaf(w.z, t[o], r.z); r.w = fmaf(w.w, t[o], r.w);
  }
  const float4 pp = *reinterpret_cast<const float4 *>(pe + (size_t)pos * kD + c4 * 4);
  *reinterpret_cast<float4 *>(out + row * kD + c4 * 4) = make_float4(r.x + pp.x, r.y + pp.y, r.z + pp.z, r.w + pp.w);
}

// DistillLayer conv input (customized_transformer.py:21-25,32): row (b, t) -> [x[t-1] | x[t] | x[t+1]] with circular wrap
__global__ void __launch_bounds__(384) mtio_im2col_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t rows, int T) {
  const int64_t row = blockIdx.x;          // one block of 384 threads per row: thread = (tap, 16-byte column piece)
  const int rem = (int)threadIdx.x, k = rem >> 7, c4 = rem & 127;
  if (row >= rows) return;
  const int64_t b = (uint32_t)blockIdx.x / (uint32_t)T;
  const int t = (int)((uint32_t)blockIdx.x % (uint32_t)T);
  const int src_t = (t + k - 1 + T) % T;
  *reinterpret_cast<float4 *>(out + row * (3 * kD) + k * kD + c4 * 4) =
      *reinterpret_cast<const float4 *>(in + (b * T + src_t) * kD + c4 * 4);
}

// MaxPool1d(kernel 3, stride 2, padding 1) over the T tokens of a sample (customized_transformer.py:28,35)
__global__ void __launch_bounds__(128) mtio_maxpool_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t n, int T, int To) {
  const int64_t orow = blockIdx.x;         // one block of 128 threads per output row
  const int c4 = (int)threadIdx.x;
  if (orow >= n * To) return;
  const int64_t b = (uint32_t)blockIdx.x / (uint32_t)To;
  const int o = (int)((uint32_t)blockIdx.x % (uint32_t)To);
  const int lo = max(2 * o - 1, 0), hi = min(2 * o + 1, T - 1);
  float4 m = *reinterpret_cast<const float4 *>(in + (b * T + lo) * kD + c4 * 4);
  for (int t = lo + 1; t <= hi; ++t) {
    const float4 v = *reinterpret_cast<const float4 *>(in + (b * T + t) * kD + c4 * 4);
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
  }
  *reinterpret_cast<float4 *>(out + orow * kD + c4 * 4) = m;
}

// nn.MultiheadAttention core for short sequences: one block per sample, one warp per head.
struct AttnArgs {
  const float *q; int64_t q_ld; int32_t Tq;          // query rows b * Tq + i
  const float *k, *v; int64_t kv_ld; int32_t kv_rows, Tk;   // key rows b * kv_rows + j, j < Tk <= 32
  float *out; int64_t out_ld;
};
// Lane = (key parity g = lane >> 4, 16-byte part of the 64-float head slice): a half-warp reads one key / value row
// slice as one 256-byte run, so every sector fetched is fully used; key 2 * it + g is handled in iteration `it`.
template <int NIT>
__global__ void __launch_bounds__(32 * kMtioHeads) mtio_attn_kernel(const AttnArgs a) {
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 4, part = lane & 15;
  const int64_t b = blockIdx.x;
  const float *kbase = a.k + b * a.kv_rows * a.kv_ld + head * kDh + part * 4;
  const float *vbase = a.v + b * a.kv_rows * a.kv_ld + head * kDh + part * 4;
  for (int qi = 0; qi < a.Tq; ++qi) {
    const float4 q4 = *reinterpret_cast<const float4 *>(a.q + (b * a.Tq + qi) * a.q_ld + head * kDh + part * 4);
    float4 k4[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int key = 2 * it + g;
      k4[it] = key < a.Tk ? *reinterpret_cast<const float4 *>(kbase + (int64_t)key * a.kv_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s[NIT];
    float mx = -INFINITY;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      float d = fmaf(q4.x, k4[it].x, fmaf(q4.y, k4[it].y, fmaf(q4.z, k4[it].z, q4.w * k4[it].w)));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      s[it] = (2 * it + g) < a.Tk ? d * 0.125f : -INFINITY;      // 1 / sqrt(head_dim 64)
      mx = fmaxf(mx, s[it]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      s[it] = (2 * it + g) < a.Tk ? expf(s[it] - mx) : 0.f;
      sum += s[it];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    const float inv = 1.0f / sum;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int key = 2 * it + g;
      if (key < a.Tk) {
        const float4 v4 = *reinterpret_cast<const float4 *>(vbase + (int64_t)key * a.kv_ld);
        const float p = s[it] * inv;
        acc.x = fmaf(p, v4.x, acc.x); acc.y = fmaf(p, v4.y, acc.y); acc.z = fmaf(p, v4.z, acc.z); acc.w = fmaf(p, v4.w, acc.w);
      }
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (g == 0) *reinterpret_cast<float4 *>(a.out + (b * a.Tq + qi) * a.out_ld + head * kDh + part * 4) = acc;
  }
}

// One decoding step's tail (mtio.py:121-131): final decoder norm -> predictor -> sigmoid -> token t + 1, ensemble over
// the MTIO heads -> wrap -> pred[b][t]; then the embedding + positional encoding of token t + 1 overwrite x[b].
struct HeadArgs {
  float *x;                       // [n][512]: decoder output of step t in, embedding of token t + 1 out
  const float *norm_w, *norm_b, *pred_w, *pred_b, *emb_w, *emb_b, *pe;
  float *tokens;                  // [n][F + 1][6]
  float *pred;                    // [n][F][2]
  int32_t n, t, F;
};
__global__ void __launch_bounds__(256) mtio_head_kernel(const HeadArgs h) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= h.n) return;
  float4 x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4 *>(h.x + b * kD + i * 128 + lane * 4);
  warp_layer_norm(x, h.norm_w, h.norm_b, lane);
  float p[kTok];
#pragma unroll
  for (int o = 0; o < kTok; ++o) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 w = *reinterpret_cast<const float4 *>(h.pred_w + o * kD + i * 128 + lane * 4);
      acc = fmaf(x[i].x, w.x, acc); acc = fmaf(x[i].y, w.y, acc); acc = fmaf(x[i].z, w.z, acc); acc = fmaf(x[i].w, w.w, acc);
    }
    const float z = warp_sum(acc) + h.pred_b[o];
    p[o] = 1.0f / (1.0f + expf(-z));
  }
  if (lane == 0) {
    float *tk = h.tokens + (b * (h.F + 1) + h.t + 1) * kTok;
#pragma unroll
    for (int o = 0; o < kTok; ++o) tk[o] = p[o];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v = ((p[c] + p[c + 2]) + p[c + 4]) / 3.0f;            // mtio.py:125-129
      if (v < 0.f) v = v - truncf(v) + 1.0f;                      // utils/common.py:61-70
      else if (v > 1.f) v = v - truncf(v);
      h.pred[(b * h.F + h.t) * 2 + c] = v;
    }
  }
  if (h.t + 1 >= h.F) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = i * 128 + lane * 4;
    float4 r = *reinterpret_cast<const float4 *>(h.emb_b + c);
#pragma unroll
    for (int o = 0; o < kTok; ++o) {    // transposed [6][512] copy
      const float4 w = *reinterpret_cast<const float4 *>(h.emb_w + o * kD + c);
      r.x = fmaf(w.x, p[o], r.x); r.y = fmaf(w.y, p[o], r.y); r.z = fmaf(w.z, p[o], r.z); r.w = fmaf(w.w, p[o], r.w);
    }
    const float4 pp = *reinterpret_cast<const float4 *>(h.pe + (size_t)(h.t + 1) * kD + c);
    *reinterpret_cast<float4 *>(h.x + b * kD + c) = make_float4(r.x + pp.x, r.y + pp.y, r.z + pp.z, r.w + pp.w);
  }
}

// LinearRegression.sample (viewport_prediction/models/linear_regression.py:16-33): per sample and coordinate an
// ordinary least-squares line through the his_window + 1 known points (x = 0, 1, ...), evaluated at the next fut_window
// positions; float64 like sklearn, stored as float32; no wrap into the unit square (the reference applies none here).
__global__ void __launch_bounds__(256) linreg_kernel(const float *__restrict__ hist, const float *__restrict__ cur, int64_t n, int T, int F,
                                                     float *__restrict__ pred) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * n) return;
  const int64_t b = idx >> 1;
  const int c = (int)(idx & 1);
  const int P = T + 1;
  const double xbar = 0.5 * (double)(P - 1);
  double ybar = 0.0;
  for (int j = 0; j < P; ++j) ybar += (double)(j < T ? hist[(b * T + j) * 2 + c] : cur[b * 2 + c]);
  ybar /= (double)P;
  double sxy = 0.0, sxx = 0.0;
  for (int j = 0; j < P; ++j) {
    const double dx = (double)j - xbar;
    const double y = (double)(j < T ? hist[(b * T + j) * 2 + c] : cur[b * 2 + c]);
    sxy += dx * (y - ybar);
    sxx += dx * dx;
  }
  const double slope = sxy / sxx, icpt = ybar - slope * xbar;
  for (int f = 0; f < F; ++f) pred[(b * F + f) * 2 + c] = (float)(icpt + slope * (double)(P + f));
}

}  // namespace mansy

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
using namespace mansy;

#define MTIO_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return mansy::set_error(MANSY_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

namespace {
struct AttnDev {
  float *w = nullptr, *b = nullptr, *ow = nullptr, *ob = nullptr;
  CUtensorMap map_qkv, map_q, map_kv, map_o;
};
struct LayerDev {
  AttnDev sa, ca;
  float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
  float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr, *n3w = nullptr, *n3b = nullptr;
  CUtensorMap map_w1, map_w2;
};
}  // namespace

struct mansy_mtio {
  int device = 0;
  int n_enc = 0, n_dec = 0, T = 0, F = 0, Tm = 0, max_batch = 0;
  bool tc_ok = false;
  int cluster_ln = 1, cluster_wide = 1;      // CTAs sharing weight boxes by TMA multicast (MANSY_MTIO_CLUSTER_LN / _WIDE)
  int w_box_rows = 128;                      // rows per weight TMA operation (MANSY_MTIO_WBOX = 128 | 256)
  int timeline_launch = -1, gemm_launches = 0, timeline_shape[4] = {0, 0, 0, 0};   // MANSY_MTIO_TIMELINE=<k>: stamp the k-th GEMM launch
  long long *timeline_dev = nullptr;
  int lanes = 2;                             // halves of a pass run on two streams (MANSY_MTIO_LANES = 1 | 2)
  int ln_split = 1;                          // LayerNorm GEMMs as 2-CTA clusters of [128 x 256] halves, row statistics swapped over
                                             // DSMEM (MANSY_MTIO_LN_SPLIT=0: one [128 x 512] CTA per row tile)
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  std::vector<void *> allocs;
  LayerDev enc[MANSY_MTIO_MAX_LAYERS], dec[MANSY_MTIO_MAX_LAYERS];
  float *emb_w = nullptr, *emb_b = nullptr, *pe = nullptr;
  float *encn_w = nullptr, *encn_b = nullptr, *decn_w = nullptr, *decn_b = nullptr;
  float *conv_w = nullptr, *conv_b = nullptr;        // [512][3 * 512] (tap-major K), BatchNorm folded in
  float *pred_w = nullptr, *pred_b = nullptr;
  CUtensorMap map_conv;
  // workspace (sized for max_batch samples)
  float *xs = nullptr, *att_e = nullptr, *x1_e = nullptr, *ff_e = nullptr, *wide_e = nullptr;   // encoder rows; wide_e = qkv / im2col [rows][1536]
  float *mem = nullptr, *memkv[MANSY_MTIO_MAX_LAYERS] = {};
  float *x = nullptr, *q = nullptr, *att = nullptr, *x1 = nullptr, *x2 = nullptr, *ffh = nullptr;
  float *kc[MANSY_MTIO_MAX_LAYERS] = {}, *vc[MANSY_MTIO_MAX_LAYERS] = {};
  float *tokens = nullptr;
  float *io_hist = nullptr, *io_cur = nullptr, *io_pred = nullptr;       // staging of the *_host call
  // timing
  bool timed = false;
  std::vector<cudaEvent_t> events;
  std::vector<int> ev_class;
  size_t ev_used = 0;
};

namespace {

int dev_alloc(mansy_mtio *m, float **p, size_t floats) {
  void *d = nullptr;
  cudaError_t e = cudaMalloc(&d, floats * sizeof(float));
  if (e != cudaSuccess) return set_error(MANSY_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  m->allocs.push_back(d);
  *p = static_cast<float *>(d);
  return MANSY_OK;
}

int upload(mansy_mtio *m, float **p, const float *host, size_t floats) {   // NULL host = zeros
  int rc = dev_alloc(m, p, floats);
  if (rc) return rc;
  if (host) MTIO_CUDA(cudaMemcpy(*p, host, floats * sizeof(float), cudaMemcpyHostToDevice));
  else MTIO_CUDA(cudaMemset(*p, 0, floats * sizeof(float)));
  return MANSY_OK;
}

int upload_attn(mansy_mtio *m, AttnDev *d, const mansy_mtio_attn_t *a, const char *what) {
  if (!a->in_proj_w || !a->out_w) return set_error(MANSY_E_INVALID, std::string(what) + ": NULL attention weight");
  int rc = upload(m, &d->w, a->in_proj_w, (size_t)3 * kD * kD);
  if (!rc) rc = upload(m, &d->b, a->in_proj_b, (size_t)3 * kD);
  if (!rc) rc = upload(m, &d->ow, a->out_w, (size_t)kD * kD);
  if (!rc) rc = upload(m, &d->ob, a->out_b, kD);
  return rc;
}

int make_attn_maps(AttnDev *d, uint32_t box_rows) {
  int rc = tc_make_map(&d->map_qkv, d->w, kD, 3 * kD, kD, box_rows);
  if (!rc) rc = tc_make_map(&d->map_q, d->w, kD, kD, kD, box_rows);
  if (!rc) rc = tc_make_map(&d->map_kv, d->w + (size_t)kD * kD, kD, 2 * kD, kD, box_rows);
  if (!rc) rc = tc_make_map(&d->map_o, d->ow, kD, kD, kD, box_rows);
  return rc;
}

// ---- launches ----
struct Launcher {
  mansy_mtio *m;
  cudaStream_t s;
  bool fp32;
  bool timed;
  int rc = MANSY_OK;

  bool begin(int cls) {
    if (rc) return false;
    if (timed) {
      if (m->ev_used + 2 > m->events.size()) {
        for (int i = 0; i < 64; ++i) {
          cudaEvent_t e;
          if (cudaEventCreate(&e) != cudaSuccess) { rc = set_error(MANSY_E_CUDA, "cudaEventCreate failed"); return false; }
          m->events.push_back(e);
        }
      }
      m->ev_class.push_back(cls);
      cudaEventRecord(m->events[m->ev_used++], s);
    }
    return true;
  }
  void end(const char *what) {
    if (timed) cudaEventRecord(m->events[m->ev_used++], s);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && !rc) rc = set_error(MANSY_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }

  template <int BN, int MT, int EPI>
  void tc_gemm(const CUtensorMap &ma, const CUtensorMap &mw, const CUtensorMap &mres, const OutMaps &mout, const GemmArgs &g) {
    using Cfg = GemmCfg<BN, MT>;
    unsigned tiles = (unsigned)((g.M + 128 * MT - 1) / (128 * MT));
    int cl = EPI == EPI_LN ? m->cluster_ln : m->cluster_wide;
    if (cl > (int)tiles) cl = 1;
    if (Cfg::kCtasPerSm > 1) cl = 1;
    tiles = (tiles + cl - 1) / cl * cl;              // whole clusters: surplus CTAs run on zero-filled / clipped rows
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles, (unsigned)(g.N / BN), 1);
    if (EPI == EPI_LN2) {                            // 2-CTA clusters along x: the two column halves of one row tile
      cl = 2;
      cfg.gridDim = dim3(2 * tiles, 1, 1);
    }
    cfg.blockDim = dim3(Cfg::kThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cl > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mtio_gemm_kernel<BN, MT, EPI>, ma, mw, mres, mout, g);
    if (e != cudaSuccess && !rc) rc = set_error(MANSY_E_CUDA, std::string("mtio_gemm_kernel launch: ") + cudaGetErrorString(e));
  }

  // C[M][N] = epi(A[M][K] W[N][K]^T + bias): `wmap` / `W` describe the same torch-layout weight rows
  void gemm(int epi, const float *A, int64_t lda, const CUtensorMap &wmap, const float *W, GemmArgs g) {
    if (!begin(0)) return;
    if (fp32) {
      dim3 grid((unsigned)((g.M + 63) / 64), (unsigned)(g.N / 64), 1);
      switch (epi) {
        case EPI_NONE: mtio_sgemm_kernel<EPI_NONE><<<grid, 256, 0, s>>>(A, lda, W, g); break;
        case EPI_RELU: mtio_sgemm_kernel<EPI_RELU><<<grid, 256, 0, s>>>(A, lda, W, g); break;
        case EPI_ELU: mtio_sgemm_kernel<EPI_ELU><<<grid, 256, 0, s>>>(A, lda, W, g); break;
        default: mtio_sgemm_kernel<EPI_LN><<<grid, 256, 0, s>>>(A, lda, W, g); break;
      }
      end("mtio_sgemm_kernel");
      if (epi == EPI_LN && begin(2)) {
        mtio_ln_kernel<<<(unsigned)((g.M + 7) / 8), 256, 0, s>>>(g.out[0], g.out[0], g.gamma, g.beta, g.M);
        end("mtio_ln_kernel");
      }
      return;
    }
    g.w_box_rows = m->w_box_rows;
    g.timeline = nullptr;
    if (m->timeline_launch >= 0 && m->gemm_launches++ == m->timeline_launch && m->timeline_dev) {
      g.timeline = m->timeline_dev;
      m->timeline_shape[0] = g.M; m->timeline_shape[1] = g.N; m->timeline_shape[2] = g.K; m->timeline_shape[3] = epi;
    }
    CUtensorMap ma;
    if (int e = tc_make_map(&ma, A, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)lda, 128)) { rc = e; return; }
    CUtensorMap mres = ma;                 // only the LayerNorm epilogue reads it
    OutMaps mout;
    for (int sgm = 0; sgm < 3; ++sgm) {
      mout.m[sgm] = ma;
      if (sgm * kD < g.N)
        if (int e = tc_make_map(&mout.m[sgm], g.out[sgm], kD, (uint64_t)g.M, (uint64_t)g.out_ld[sgm], 128)) { rc = e; return; }
    }
    if (epi == EPI_LN)
      if (int e = tc_make_map(&mres, g.res, kD, (uint64_t)g.M, (uint64_t)g.res_ld, 128)) { rc = e; return; }
    // two M128 sub-tiles per CTA share the weight boxes -- when that still leaves enough CTAs to occupy the SMs; small
    // batches (e.g. 2 048 samples per GPU in BASELINE config 5) keep [128 x 256] tiles, two CTAs per SM
    static const int wide_min = getenv("MANSY_MTIO_WIDE_MIN_CTAS") ? atoi(getenv("MANSY_MTIO_WIDE_MIN_CTAS")) : 64;
    const bool wide = g.M > 128 && ((g.M + 255) / 256) * (g.N / 256) >= wide_min;
    switch (epi) {
      case EPI_NONE: wide ? tc_gemm<256, 2, EPI_NONE>(ma, wmap, mres, mout, g) : tc_gemm<256, 1, EPI_NONE>(ma, wmap, mres, mout, g); break;
      case EPI_RELU: wide ? tc_gemm<256, 2, EPI_RELU>(ma, wmap, mres, mout, g) : tc_gemm<256, 1, EPI_RELU>(ma, wmap, mres, mout, g); break;
      case EPI_ELU: wide ? tc_gemm<256, 2, EPI_ELU>(ma, wmap, mres, mout, g) : tc_gemm<256, 1, EPI_ELU>(ma, wmap, mres, mout, g); break;
      default:
        if (m->ln_split) tc_gemm<256, 1, EPI_LN2>(ma, wmap, mres, mout, g);
        else tc_gemm<512, 1, EPI_LN>(ma, wmap, mres, mout, g);
        break;
    }
    end("mtio_gemm_kernel");
  }

  void attn(const AttnArgs &a, int64_t n) {
    if (!begin(1)) return;
    const int nit = (a.Tk + 1) / 2;
    if (nit <= 2) mtio_attn_kernel<2><<<(unsigned)n, 32 * kMtioHeads, 0, s>>>(a);
    else if (nit <= 4) mtio_attn_kernel<4><<<(unsigned)n, 32 * kMtioHeads, 0, s>>>(a);
    else if (nit <= 8) mtio_attn_kernel<8><<<(unsigned)n, 32 * kMtioHeads, 0, s>>>(a);
    else mtio_attn_kernel<16><<<(unsigned)n, 32 * kMtioHeads, 0, s>>>(a);
    end("mtio_attn_kernel");
  }
};

template <int BN, int MT, int EPI>
void gemm_set_attribute() {
  cudaFuncSetAttribute(mtio_gemm_kernel<BN, MT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<BN, MT>::kSmem);
}
void gemm_set_attributes() {
  gemm_set_attribute<256, 1, EPI_NONE>(); gemm_set_attribute<256, 1, EPI_RELU>(); gemm_set_attribute<256, 1, EPI_ELU>();
  gemm_set_attribute<256, 2, EPI_NONE>(); gemm_set_attribute<256, 2, EPI_RELU>(); gemm_set_attribute<256, 2, EPI_ELU>();
  gemm_set_attribute<512, 1, EPI_LN>();
  gemm_set_attribute<256, 1, EPI_LN2>();
  cudaGetLastError();
}

GemmArgs gemm_args(int M, int N, int K, const float *bias) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.bias = bias;
  return g;
}

// Workspace of one lane: the per-sample buffers of the handle, offset by `first` samples.
struct Ws {
  float *xs, *att_e, *x1_e, *ff_e, *wide_e, *mem, *memkv[MANSY_MTIO_MAX_LAYERS];
  float *x, *q, *att, *x1, *x2, *ffh, *kc[MANSY_MTIO_MAX_LAYERS], *vc[MANSY_MTIO_MAX_LAYERS], *tokens;
};
Ws workspace(const mansy_mtio *m, size_t first) {
  Ws w;
  const size_t T = (size_t)m->T, F = (size_t)m->F, Tm = (size_t)m->Tm;
  w.xs = m->xs + first * T * kD; w.att_e = m->att_e + first * T * kD; w.x1_e = m->x1_e + first * T * kD;
  w.ff_e = m->ff_e + first * T * kD; w.wide_e = m->wide_e + first * T * 3 * kD; w.mem = m->mem + first * Tm * kD;
  for (int l = 0; l < MANSY_MTIO_MAX_LAYERS; ++l) {
    w.memkv[l] = m->memkv[l] ? m->memkv[l] + first * Tm * 2 * kD : nullptr;
    w.kc[l] = m->kc[l] ? m->kc[l] + first * F * kD : nullptr;
    w.vc[l] = m->vc[l] ? m->vc[l] + first * F * kD : nullptr;
  }
  w.x = m->x + first * kD; w.q = m->q + first * kD; w.att = m->att + first * kD; w.x1 = m->x1 + first * kD;
  w.x2 = m->x2 + first * kD; w.ffh = m->ffh + first * kD; w.tokens = m->tokens + first * (F + 1) * kTok;
  return w;
}

// one pass over n <= max_batch samples
int run_chunk(mansy_mtio *m, const Ws &w, const float *hist, const float *cur, int n, int n_steps, int flags, float *pred,
              float *tokens_out, cudaStream_t s, bool timed) {
  Launcher L{m, s, (flags & MANSY_MTIO_FP32) != 0, timed};
  if (!L.fp32 && !m->tc_ok) return set_error(MANSY_E_STATE, "tensor maps unavailable (cuTensorMapEncodeTiled missing); use MANSY_MTIO_FP32");
  const int T = m->T, F = m->F, Tm = m->Tm;
  const int rows = n * T;

  // ---- encoder (once per sample) ----
  if (L.begin(2)) {
    mtio_embed_kernel<<<(unsigned)rows, 128, 0, s>>>(hist, 2, 2, rows, T, 0, m->emb_w, m->emb_b, m->pe, w.xs, nullptr, 0);
    L.end("mtio_embed_kernel");
  }
  for (int l = 0; l < m->n_enc; ++l) {
    LayerDev &E = m->enc[l];
    GemmArgs g = gemm_args(rows, 3 * kD, kD, E.sa.b);
    for (int sgm = 0; sgm < 3; ++sgm) { g.out[sgm] = w.wide_e + sgm * kD; g.out_ld[sgm] = 3 * kD; }
    L.gemm(EPI_NONE, w.xs, kD, E.sa.map_qkv, E.sa.w, g);
    AttnArgs a{w.wide_e, 3 * kD, T, w.wide_e + kD, w.wide_e + 2 * kD, 3 * kD, T, T, w.att_e, kD};
    L.attn(a, n);
    g = gemm_args(rows, kD, kD, E.sa.ob);
    g.out[0] = w.x1_e; g.out_ld[0] = kD; g.res = w.xs; g.res_ld = kD; g.gamma = E.n1w; g.beta = E.n1b;
    L.gemm(EPI_LN, w.att_e, kD, E.sa.map_o, E.sa.ow, g);
    g = gemm_args(rows, kD, kD, E.b1);
    g.out[0] = w.ff_e; g.out_ld[0] = kD;
    L.gemm(EPI_RELU, w.x1_e, kD, E.map_w1, E.w1, g);
    g = gemm_args(rows, kD, kD, E.b2);
    g.out[0] = w.xs; g.out_ld[0] = kD; g.res = w.x1_e; g.res_ld = kD; g.gamma = E.n2w; g.beta = E.n2b;
    L.gemm(EPI_LN, w.ff_e, kD, E.map_w2, E.w2, g);
  }
  if (L.begin(2)) {   // encoder.norm, then the DistillLayer: im2col -> conv GEMM (+ folded BatchNorm, ELU) -> max-pool
    mtio_ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(w.xs, w.att_e, m->encn_w, m->encn_b, rows);
    L.end("mtio_ln_kernel");
  }
  if (L.begin(2)) {
    mtio_im2col_kernel<<<(unsigned)rows, 384, 0, s>>>(w.att_e, w.wide_e, rows, T);
    L.end("mtio_im2col_kernel");
  }
  {
    GemmArgs g = gemm_args(rows, kD, 3 * kD, m->conv_b);
    g.out[0] = w.x1_e; g.out_ld[0] = kD;
    L.gemm(EPI_ELU, w.wide_e, 3 * kD, m->map_conv, m->conv_w, g);
  }
  if (L.begin(2)) {
    mtio_maxpool_kernel<<<(unsigned)(n * Tm), 128, 0, s>>>(w.x1_e, w.mem, n, T, Tm);
    L.end("mtio_maxpool_kernel");
  }
  for (int l = 0; l < m->n_dec; ++l) {   // cross-attention keys / values of the memory, once per layer
    AttnDev &C = m->dec[l].ca;
    GemmArgs g = gemm_args(n * Tm, 2 * kD, kD, C.b + kD);
    g.out[0] = w.memkv[l]; g.out_ld[0] = 2 * kD;
    g.out[1] = w.memkv[l] + kD; g.out_ld[1] = 2 * kD;
    L.gemm(EPI_NONE, w.mem, kD, C.map_kv, C.w + (size_t)kD * kD, g);
  }

  // ---- decoder: F autoregressive steps, one new token per step ----
  if (L.begin(2)) {
    mtio_embed_kernel<<<(unsigned)n, 128, 0, s>>>(cur, 2, 2, n, 1, 0, m->emb_w, m->emb_b, m->pe, w.x, w.tokens, (int64_t)(F + 1) * kTok);
    L.end("mtio_embed_kernel");
  }
  for (int t = 0; t < n_steps; ++t) {
    for (int l = 0; l < m->n_dec; ++l) {
      LayerDev &D = m->dec[l];
      GemmArgs g = gemm_args(n, 3 * kD, kD, D.sa.b);
      g.out[0] = w.q; g.out_ld[0] = kD;
      g.out[1] = w.kc[l] + (size_t)t * kD; g.out_ld[1] = (int64_t)F * kD;
      g.out[2] = w.vc[l] + (size_t)t * kD; g.out_ld[2] = (int64_t)F * kD;
      L.gemm(EPI_NONE, w.x, kD, D.sa.map_qkv, D.sa.w, g);
      AttnArgs a{w.q, kD, 1, w.kc[l], w.vc[l], kD, F, t + 1, w.att, kD};
      L.attn(a, n);
      g = gemm_args(n, kD, kD, D.sa.ob);
      g.out[0] = w.x1; g.out_ld[0] = kD; g.res = w.x; g.res_ld = kD; g.gamma = D.n1w; g.beta = D.n1b;
      L.gemm(EPI_LN, w.att, kD, D.sa.map_o, D.sa.ow, g);
      g = gemm_args(n, kD, kD, D.ca.b);
      g.out[0] = w.q; g.out_ld[0] = kD;
      L.gemm(EPI_NONE, w.x1, kD, D.ca.map_q, D.ca.w, g);
      AttnArgs c{w.q, kD, 1, w.memkv[l], w.memkv[l] + kD, 2 * kD, Tm, Tm, w.att, kD};
      L.attn(c, n);
      g = gemm_args(n, kD, kD, D.ca.ob);
      g.out[0] = w.x2; g.out_ld[0] = kD; g.res = w.x1; g.res_ld = kD; g.gamma = D.n2w; g.beta = D.n2b;
      L.gemm(EPI_LN, w.att, kD, D.ca.map_o, D.ca.ow, g);
      g = gemm_args(n, kD, kD, D.b1);
      g.out[0] = w.ffh; g.out_ld[0] = kD;
      L.gemm(EPI_RELU, w.x2, kD, D.map_w1, D.w1, g);
      g = gemm_args(n, kD, kD, D.b2);
      g.out[0] = w.x; g.out_ld[0] = kD; g.res = w.x2; g.res_ld = kD; g.gamma = D.n3w; g.beta = D.n3b;
      L.gemm(EPI_LN, w.ffh, kD, D.map_w2, D.w2, g);
    }
    if (L.begin(2)) {
      HeadArgs h{w.x, m->decn_w, m->decn_b, m->pred_w, m->pred_b, m->emb_w, m->emb_b, m->pe, w.tokens, pred, n, t, F};
      mtio_head_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(h);
      L.end("mtio_head_kernel");
    }
  }
  if (L.rc) return L.rc;
  if (tokens_out)     // tokens 0 .. n_steps of every sample; later rows of the caller's buffer stay untouched
    MTIO_CUDA(cudaMemcpy2DAsync(tokens_out, (size_t)(F + 1) * kTok * sizeof(float), w.tokens, (size_t)(F + 1) * kTok * sizeof(float),
                                (size_t)(n_steps + 1) * kTok * sizeof(float), (size_t)n, cudaMemcpyDeviceToDevice, s));
  return MANSY_OK;
}

// A pass over n samples as two lanes on two streams: attention is bound by HBM (fp32 key/value cache) and leaves the SMs'
// shared memory and tensor memory alone, the GEMMs hold one CTA per SM, so work of one half can fill what the other
// leaves idle (partial waves, a GEMM CTA's epilogue).  Measured gain 4.7 % (15.15 -> 14.43 ms for 16 384 samples); it does
// not depend on staggering the lanes or on the split ratio, i.e. the kernels mostly still run one after the other: an
// attention CTA (20 K registers) rarely fits beside a GEMM CTA (35 K), and forcing the max-shared carve-out on the
// attention kernels so that they may share an SM costs them their L1 (15.1 ms).  The halves are cut at a 256-row tile
// boundary and launched by two host threads (the second lane's stream is ordered after the caller's stream at entry and
// joined to it at exit).
int run_lanes(mansy_mtio *m, const float *hist, const float *cur, int n, int n_steps, int flags, float *pred, float *tokens_out,
              cudaStream_t s) {
  const bool timed = (flags & MANSY_MTIO_TIME_KERNELS) != 0;
  if (m->lanes < 2 || timed || n < 1024 || !m->s2)
    return run_chunk(m, workspace(m, 0), hist, cur, n, n_steps, flags, pred, tokens_out, s, timed);
  const int h = ((n / 2 + 255) / 256) * 256;
  MTIO_CUDA(cudaEventRecord(m->ev_in, s));
  MTIO_CUDA(cudaStreamWaitEvent(m->s2, m->ev_in, 0));
  int rc1 = MANSY_OK;
  std::string msg1;
  const int F = m->F, T = m->T, device = m->device;
  std::thread lane1([&] {
    cudaSetDevice(device);
    rc1 = run_chunk(m, workspace(m, (size_t)h), hist + (size_t)h * T * 2, cur + (size_t)h * 2, n - h, n_steps, flags,
                    pred + (size_t)h * F * 2, tokens_out ? tokens_out + (size_t)h * (F + 1) * kTok : nullptr, m->s2, false);
    if (rc1) msg1 = mansy_last_error();
  });
  const int rc0 = run_chunk(m, workspace(m, 0), hist, cur, h, n_steps, flags, pred, tokens_out, s, false);
  lane1.join();
  MTIO_CUDA(cudaEventRecord(m->ev_out, m->s2));
  MTIO_CUDA(cudaStreamWaitEvent(s, m->ev_out, 0));
  if (rc0) return rc0;
  if (rc1) return set_error(rc1, msg1);
  return MANSY_OK;
}

}  // namespace

extern "C" {

int mansy_mtio_destroy(mansy_mtio_t m) {
  if (!m) return MANSY_OK;
  DeviceScope dscope(m->device);
  for (void *p : m->allocs) cudaFree(p);
  for (cudaEvent_t e : m->events) cudaEventDestroy(e);
  if (m->s2) cudaStreamDestroy(m->s2);
  if (m->ev_in) cudaEventDestroy(m->ev_in);
  if (m->ev_out) cudaEventDestroy(m->ev_out);
  delete m;
  return MANSY_OK;
}

int mansy_mtio_create(const mansy_mtio_weights_t *w, int device, int32_t max_batch, mansy_mtio_t *out) {
  if (!w || !out) return set_error(MANSY_E_INVALID, "NULL argument");
  *out = nullptr;
  if (w->n_enc < 1 || w->n_enc > MANSY_MTIO_MAX_LAYERS || w->n_dec < 1 || w->n_dec > MANSY_MTIO_MAX_LAYERS)
    return set_error(MANSY_E_INVALID, "n_enc / n_dec must be 1..4");
  if (w->his_window < 1 || w->his_window > 16 || w->fut_window < 1 || w->fut_window > 31)
    return set_error(MANSY_E_INVALID, "his_window must be 1..16 and fut_window 1..31");
  if (w->pe_rows < w->his_window || w->pe_rows < w->fut_window) return set_error(MANSY_E_INVALID, "pe_rows too small");
  if (max_batch < 1) return set_error(MANSY_E_INVALID, "max_batch must be >= 1");
  if (!w->emb_w || !w->emb_b || !w->pe || !w->enc_norm_w || !w->dec_norm_w || !w->conv_w || !w->conv_b || !w->bn_w || !w->bn_b ||
      !w->bn_mean || !w->bn_var || !w->pred_w || !w->pred_b)
    return set_error(MANSY_E_INVALID, "a required weight pointer is NULL");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
    cudaGetLastError();
    return set_error(MANSY_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
  }
  DeviceScope dscope(device);
  MTIO_CUDA(dscope.err);
  mansy_mtio *m = new (std::nothrow) mansy_mtio();
  if (!m) return set_error(MANSY_E_NOMEM, "out of host memory");
  m->device = device;
  m->n_enc = w->n_enc; m->n_dec = w->n_dec; m->T = w->his_window; m->F = w->fut_window;
  m->Tm = (m->T + 2 - 3) / 2 + 1;
  m->max_batch = max_batch;
  if (const char *v = getenv("MANSY_MTIO_CLUSTER_LN")) m->cluster_ln = atoi(v) == 4 ? 4 : (atoi(v) == 2 ? 2 : 1);
  if (const char *v = getenv("MANSY_MTIO_CLUSTER_WIDE")) m->cluster_wide = atoi(v) == 2 ? 2 : 1;
  if (const char *v = getenv("MANSY_MTIO_WBOX")) m->w_box_rows = atoi(v) == 256 ? 256 : 128;
  if (const char *v = getenv("MANSY_MTIO_LANES")) m->lanes = atoi(v) == 1 ? 1 : 2;
  if (const char *v = getenv("MANSY_MTIO_TIMELINE")) {
    m->timeline_launch = atoi(v);
    m->lanes = 1;
    void *d = nullptr;
    if (cudaMalloc(&d, 256 * sizeof(long long)) == cudaSuccess) {
      cudaMemset(d, 0, 256 * sizeof(long long));
      m->allocs.push_back(d);
      m->timeline_dev = static_cast<long long *>(d);
    }
  }
  if (const char *v = getenv("MANSY_MTIO_LN_SPLIT")) m->ln_split = atoi(v) != 0;
  int rc = MANSY_OK;
#define MTIO_TRY(expr) do { if (!rc) rc = (expr); } while (0)
  {   // embedding.linear.weight [512][6] -> [6][512]: a warp reads 32 consecutive outputs of one input column
    std::vector<float> wt((size_t)kTok * kD);
    for (int c = 0; c < kD; ++c)
      for (int o = 0; o < kTok; ++o) wt[(size_t)o * kD + c] = w->emb_w[(size_t)c * kTok + o];
    MTIO_TRY(upload(m, &m->emb_w, wt.data(), wt.size()));
  }
  MTIO_TRY(upload(m, &m->emb_b, w->emb_b, kD));
  MTIO_TRY(upload(m, &m->pe, w->pe, (size_t)w->pe_rows * kD));
  MTIO_TRY(upload(m, &m->encn_w, w->enc_norm_w, kD));
  MTIO_TRY(upload(m, &m->encn_b, w->enc_norm_b, kD));
  MTIO_TRY(upload(m, &m->decn_w, w->dec_norm_w, kD));
  MTIO_TRY(upload(m, &m->decn_b, w->dec_norm_b, kD));
  MTIO_TRY(upload(m, &m->pred_w, w->pred_w, (size_t)kTok * kD));
  MTIO_TRY(upload(m, &m->pred_b, w->pred_b, kTok));
  for (int l = 0; l < m->n_enc + m->n_dec && !rc; ++l) {
    const bool is_dec = l >= m->n_enc;
    const mansy_mtio_layer_t *h = is_dec ? &w->dec[l - m->n_enc] : &w->enc[l];
    LayerDev *d = is_dec ? &m->dec[l - m->n_enc] : &m->enc[l];
    if (!h->lin1_w || !h->lin2_w || !h->norm1_w || !h->norm2_w || (is_dec && !h->norm3_w)) {
      rc = set_error(MANSY_E_INVALID, "a required layer weight pointer is NULL");
      break;
    }
    MTIO_TRY(upload_attn(m, &d->sa, &h->self_attn, "self_attn"));
    if (is_dec) MTIO_TRY(upload_attn(m, &d->ca, &h->cross_attn, "multihead_attn"));
    MTIO_TRY(upload(m, &d->w1, h->lin1_w, (size_t)kD * kD));
    MTIO_TRY(upload(m, &d->b1, h->lin1_b, kD));
    MTIO_TRY(upload(m, &d->w2, h->lin2_w, (size_t)kD * kD));
    MTIO_TRY(upload(m, &d->b2, h->lin2_b, kD));
    MTIO_TRY(upload(m, &d->n1w, h->norm1_w, kD));
    MTIO_TRY(upload(m, &d->n1b, h->norm1_b, kD));
    MTIO_TRY(upload(m, &d->n2w, h->norm2_w, kD));
    MTIO_TRY(upload(m, &d->n2b, h->norm2_b, kD));
    if (is_dec) {
      MTIO_TRY(upload(m, &d->n3w, h->norm3_w, kD));
      MTIO_TRY(upload(m, &d->n3b, h->norm3_b, kD));
    }
  }
  if (!rc) {
    // DistillLayer: y = ELU(BN(conv(x))) with BN in eval mode = per-channel affine, folded into the conv:
    // W'[o][k * 512 + i] = W[o][i][k] * g[o] / sqrt(var[o] + eps),  b'[o] = (b[o] - mean[o]) * g[o] / sqrt(var[o] + eps) + beta[o]
    std::vector<float> cw((size_t)kD * 3 * kD), cb(kD);
    for (int o = 0; o < kD; ++o) {
      const float sc = w->bn_w[o] / std::sqrt(w->bn_var[o] + 1e-5f);
      for (int i = 0; i < kD; ++i)
        for (int k = 0; k < 3; ++k) cw[(size_t)o * 3 * kD + (size_t)k * kD + i] = w->conv_w[((size_t)o * kD + i) * 3 + k] * sc;
      cb[o] = (w->conv_b[o] - w->bn_mean[o]) * sc + w->bn_b[o];
    }
    MTIO_TRY(upload(m, &m->conv_w, cw.data(), cw.size()));
    MTIO_TRY(upload(m, &m->conv_b, cb.data(), cb.size()));
  }
  const size_t B = (size_t)max_batch, rowsE = B * m->T;
  MTIO_TRY(dev_alloc(m, &m->xs, rowsE * kD));
  MTIO_TRY(dev_alloc(m, &m->att_e, rowsE * kD));
  MTIO_TRY(dev_alloc(m, &m->x1_e, rowsE * kD));
  MTIO_TRY(dev_alloc(m, &m->ff_e, rowsE * kD));
  MTIO_TRY(dev_alloc(m, &m->wide_e, rowsE * 3 * kD));
  MTIO_TRY(dev_alloc(m, &m->mem, B * m->Tm * kD));
  for (int l = 0; l < m->n_dec; ++l) {
    MTIO_TRY(dev_alloc(m, &m->memkv[l], B * m->Tm * 2 * kD));
    MTIO_TRY(dev_alloc(m, &m->kc[l], B * m->F * kD));
    MTIO_TRY(dev_alloc(m, &m->vc[l], B * m->F * kD));
  }
  MTIO_TRY(dev_alloc(m, &m->x, B * kD));
  MTIO_TRY(dev_alloc(m, &m->q, B * kD));
  MTIO_TRY(dev_alloc(m, &m->att, B * kD));
  MTIO_TRY(dev_alloc(m, &m->x1, B * kD));
  MTIO_TRY(dev_alloc(m, &m->x2, B * kD));
  MTIO_TRY(dev_alloc(m, &m->ffh, B * kD));
  MTIO_TRY(dev_alloc(m, &m->tokens, B * (m->F + 1) * kTok));
#undef MTIO_TRY
  if (rc) { mansy_mtio_destroy(m); return rc; }
  // tensor maps of the weights (the fp32 path works without them)
  int mrc = MANSY_OK;
  const uint32_t wb = (uint32_t)m->w_box_rows;
  for (int l = 0; l < m->n_enc && !mrc; ++l) {
    mrc = make_attn_maps(&m->enc[l].sa, wb);
    if (!mrc) mrc = tc_make_map(&m->enc[l].map_w1, m->enc[l].w1, kD, kD, kD, wb);
    if (!mrc) mrc = tc_make_map(&m->enc[l].map_w2, m->enc[l].w2, kD, kD, kD, wb);
  }
  for (int l = 0; l < m->n_dec && !mrc; ++l) {
    mrc = make_attn_maps(&m->dec[l].sa, wb);
    if (!mrc) mrc = make_attn_maps(&m->dec[l].ca, wb);
    if (!mrc) mrc = tc_make_map(&m->dec[l].map_w1, m->dec[l].w1, kD, kD, kD, wb);
    if (!mrc) mrc = tc_make_map(&m->dec[l].map_w2, m->dec[l].w2, kD, kD, kD, wb);
  }
  if (!mrc) mrc = tc_make_map(&m->map_conv, m->conv_w, 3 * kD, kD, 3 * kD, wb);
  m->tc_ok = (mrc == MANSY_OK);
  if (cudaStreamCreateWithFlags(&m->s2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_out, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    m->lanes = 1;
  }
  gemm_set_attributes();       // opt-in shared-memory sizes, once, before any lane thread launches
  *out = m;
  return MANSY_OK;
}

int mansy_mtio_sample(mansy_mtio_t m, const float *history_dev, const float *current_dev, int32_t n, int32_t n_steps, int32_t flags,
                      float *pred_dev, float *tokens_dev, void *stream) {
  if (!m || !history_dev || !current_dev || !pred_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (n < 0) return set_error(MANSY_E_INVALID, "n must be >= 0");
  if (n_steps < 0 || n_steps > m->F) return set_error(MANSY_E_INVALID, "n_steps must be 0 (= fut_window) .. fut_window");
  if (n_steps == 0) n_steps = m->F;
  DeviceScope dscope(m->device);               // the handle's device whatever the caller's current one is
  MTIO_CUDA(dscope.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  m->timed = (flags & MANSY_MTIO_TIME_KERNELS) != 0;
  m->ev_used = 0;
  m->ev_class.clear();
  for (int64_t off = 0; off < n; off += m->max_batch) {
    const int c = (int)((n - off) < m->max_batch ? (n - off) : m->max_batch);
    int rc = run_lanes(m, history_dev + off * m->T * 2, current_dev + off * 2, c, n_steps, flags, pred_dev + off * m->F * 2,
                       tokens_dev ? tokens_dev + off * (m->F + 1) * kTok : nullptr, s);
    if (rc) return rc;
  }
  if (m->timeline_dev && m->gemm_launches > m->timeline_launch && m->timeline_launch >= 0) {
    long long tl[256];
    cudaStreamSynchronize(s);
    cudaMemcpy(tl, m->timeline_dev, sizeof(tl), cudaMemcpyDeviceToHost);
    const long long t0 = tl[190];
    fprintf(stderr, "mtio_gemm timeline: launch %d  M %d N %d K %d epi %d  (SM cycles after the prologue of CTA (0,0))\n",
            m->timeline_launch, m->timeline_shape[0], m->timeline_shape[1], m->timeline_shape[2], m->timeline_shape[3]);
    for (int it = 0; it < 60 && tl[120 + it]; ++it)
      fprintf(stderr, "  chunk %2d: stage free / TMA issue %7lld   operands landed %7lld   MMAs issued %7lld\n", it, tl[it] - t0,
              tl[60 + it] - t0, tl[120 + it] - t0);
    fprintf(stderr, "  accumulators complete %lld   epilogue done %lld\n", tl[191] - t0, tl[192] - t0);
    m->timeline_launch = -1;
  }
  return MANSY_OK;
}

int mansy_mtio_sample_host(mansy_mtio_t m, const float *history_host, const float *current_host, int32_t n, int32_t n_steps,
                           int32_t flags, float *pred_host, void *stream) {
  if (!m || !history_host || !current_host || !pred_host) return set_error(MANSY_E_INVALID, "NULL argument");
  if (n < 0) return set_error(MANSY_E_INVALID, "n must be >= 0");
  if (n_steps < 0 || n_steps > m->F) return set_error(MANSY_E_INVALID, "n_steps must be 0 (= fut_window) .. fut_window");
  if (n_steps == 0) n_steps = m->F;
  DeviceScope dscope(m->device);
  MTIO_CUDA(dscope.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!m->io_hist) {
    const size_t B = (size_t)m->max_batch;
    int rc = dev_alloc(m, &m->io_hist, B * m->T * 2);
    if (!rc) rc = dev_alloc(m, &m->io_cur, B * 2);
    if (!rc) rc = dev_alloc(m, &m->io_pred, B * m->F * 2);
    if (rc) return rc;
  }
  m->timed = false;
  for (int64_t off = 0; off < n; off += m->max_batch) {
    const int c = (int)((n - off) < m->max_batch ? (n - off) : m->max_batch);
    MTIO_CUDA(cudaMemcpyAsync(m->io_hist, history_host + off * m->T * 2, (size_t)c * m->T * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    MTIO_CUDA(cudaMemcpyAsync(m->io_cur, current_host + off * 2, (size_t)c * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = run_lanes(m, m->io_hist, m->io_cur, c, n_steps, flags & ~MANSY_MTIO_TIME_KERNELS, m->io_pred, nullptr, s);
    if (rc) return rc;
    // rows 0 .. n_steps - 1 of every sample (the rest of the caller's buffer stays untouched)
    MTIO_CUDA(cudaMemcpy2DAsync(pred_host + off * m->F * 2, (size_t)m->F * 2 * sizeof(float), m->io_pred, (size_t)m->F * 2 * sizeof(float),
                                (size_t)n_steps * 2 * sizeof(float), (size_t)c, cudaMemcpyDeviceToHost, s));
  }
  MTIO_CUDA(cudaStreamSynchronize(s));
  return MANSY_OK;
}

int mansy_linreg_sample(const float *history_dev, const float *current_dev, int64_t n, int32_t his_window, int32_t fut_window,
                        float *pred_dev, void *stream) {
  if (!history_dev || !current_dev || !pred_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (n < 0 || his_window < 1 || fut_window < 1) return set_error(MANSY_E_INVALID, "n >= 0, his_window >= 1, fut_window >= 1");
  if (n == 0) return MANSY_OK;
  DeviceScope dscope(device_of_pointer(history_dev));
  linreg_kernel<<<(unsigned)((2 * n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(history_dev, current_dev, n, his_window,
                                                                                              fut_window, pred_dev);
  count_launch();
  MTIO_CUDA(cudaGetLastError());
  return MANSY_OK;
}

int mansy_mtio_kernel_ms(mansy_mtio_t m, double ms[3], int32_t launches[3]) {
  if (!m || !ms || !launches) return set_error(MANSY_E_INVALID, "NULL argument");
  for (int i = 0; i < 3; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (size_t i = 0; i < m->ev_class.size() && 2 * i + 1 < m->ev_used; ++i) {
    float t = 0.f;
    MTIO_CUDA(cudaEventElapsedTime(&t, m->events[2 * i], m->events[2 * i + 1]));
    ms[m->ev_class[i]] += t;
    launches[m->ev_class[i]] += 1;
  }
  return MANSY_OK;
}

}  // extern "C"
