// mansy_core.cuh -- scalar building blocks of the streaming simulator, shared by every kernel.
//
// Everything here is __host__ __device__ so the same code can be exercised on the CPU by the
// self-test entry points (mansy_selftest_*), which the non-GPU test-suite checks against the
// oracle before any GPU time is spent.  The kernels in mansy_sim.cu are the only product path.
//
// Numeric contract (oracle/sim_oracle.py chain="f64"): time, buffer, QoE and
// reward scalars are IEEE float64 with the reference's operation order; every operation that the
// compiler could contract into an FMA is written with a round-to-nearest intrinsic so the device
// results are bit-identical to CPython's float arithmetic.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define MANSY_HD __host__ __device__ __forceinline__
#else
#define MANSY_HD inline
#endif

namespace mansy {

// ---- exact (non-contracted) float64 / float32 arithmetic ---------------------------------
MANSY_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
MANSY_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  volatile double r = a - b; return r;
#endif
}
MANSY_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
MANSY_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b; return r;
#endif
}
// a / b, correctly rounded, from r = RN(1 / b) computed once on the host (Markstein's sequence: q = a * r refined
// twice through the exact FMA residual).  With a correctly rounded reciprocal and a faithful quotient estimate the
// last FMA rounds to RN(a / b), i.e. the result is bit-identical to an IEEE division (tests/test_capi_cpu.py checks
// 10^7 random and near-midpoint cases per run against a / b through mansy_selftest_ddiv_rcp; a standalone run of
// 6 x 10^8 cases found no mismatch).  Five dependent FMA-pipe operations instead of the ~30-instruction __ddiv_rn
// sequence: divisions were 13 % of the simulator step's instructions and sit on its critical path.
// Domain: finite a, finite b > 0 away from the over/underflow thresholds (throughputs, tile counts, rates).
MANSY_HD double ddiv_rcp(double a, double b, double r) {
#ifdef __CUDA_ARCH__
  double q = __dmul_rn(a, r);
  double e = __fma_rn(-b, q, a);
  q = __fma_rn(e, r, q);
  e = __fma_rn(-b, q, a);
  return __fma_rn(e, r, q);
#else
  volatile double q = a * r;
  volatile double e = fma(-b, q, a);
  q = fma(e, r, q);
  e = fma(-b, q, a);
  volatile double out = fma(e, r, q);
  return out;
#endif
}
MANSY_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
MANSY_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}

// ---- a1: action -> (rate_in, rate_out)  (bitrate_selection/utils/common.py:101-119) --------
// Packed 4-bit pairs; any action outside 0..14 keeps the reference's initial (0, 0).
MANSY_HD void action_to_rates(int action, int &rate_in, int &rate_out) {
  // index:            0  1  2  3  4  5  6  7  8  9 10 11 12 13 14
  // rate_in :         1  2  3  4  2  3  4  3  4  4  0  1  2  3  4
  // rate_out:         0  0  0  0  1  1  1  2  2  3  0  1  2  3  4
  const uint64_t kIn = 0x432104434324321ULL;
  const uint64_t kOut = 0x432103221110000ULL;
  if (action < 0 || action > 14) { rate_in = 0; rate_out = 0; return; }
  rate_in = (int)((kIn >> (4 * action)) & 0xF);
  rate_out = (int)((kOut >> (4 * action)) & 0xF);
}

// ---- a2: pyramid allocation (bitrate_selection/utils/common.py:142-193) ---------------------
// The reference's multi-source 8-neighbour BFS on the 8x8 torus assigns each tile its toroidal
// Chebyshev distance to the predicted viewport.  With the mask as a uint64 (bit t = tile
// row*8+col) the set of tiles within distance k is the k-fold 3x3 dilation with wrap-around,
// which is a handful of shifts and rotates.
MANSY_HD uint64_t dilate_torus(uint64_t x) {
  const uint64_t kNotCol0 = 0xFEFEFEFEFEFEFEFEULL, kCol0 = 0x0101010101010101ULL;
  const uint64_t kNotCol7 = 0x7F7F7F7F7F7F7F7FULL, kCol7 = 0x8080808080808080ULL;
  uint64_t h = x | ((x << 1) & kNotCol0) | ((x >> 7) & kCol0)   // column +1 (wraps 7 -> 0)
                 | ((x >> 1) & kNotCol7) | ((x << 7) & kCol7);  // column -1 (wraps 0 -> 7)
  return h | (h << 8) | (h >> 56) | (h >> 8) | (h << 56);        // row +-1 with wrap
}

struct TileScaleMasks {  // nested sets: d0 (viewport) c d1 c d2 c d3; everything else is scale 4
  uint64_t d0, d1, d2, d3;
};

MANSY_HD TileScaleMasks tile_scale_masks(uint64_t pred) {
  TileScaleMasks m;
  if (pred == 0) {  // empty queue: every scale stays 0 -> every tile gets rate_in (common.py:154-168)
    m.d0 = m.d1 = m.d2 = m.d3 = ~0ULL;
    return m;
  }
  m.d0 = pred;
  m.d1 = dilate_torus(m.d0);
  m.d2 = dilate_torus(m.d1);
  m.d3 = dilate_torus(m.d2);
  return m;
}

MANSY_HD int tile_scale(const TileScaleMasks &m, int tile) {
  return 4 - (int)((m.d0 >> tile) & 1) - (int)((m.d1 >> tile) & 1) - (int)((m.d2 >> tile) & 1)
           - (int)((m.d3 >> tile) & 1);
}

// lut[rate_out]: 3-bit entries, entry s (s = 1..4) = version closest to rates[rate_out] // s
// (common.py:170-190, built on the host by build_rate_lut); entry 0 is patched with rate_in.
MANSY_HD int tile_version(uint32_t lut_word, int rate_in, int scale) {
  return scale == 0 ? rate_in : (int)((lut_word >> (3 * scale)) & 7u);
}

inline void build_rate_lut(const int32_t rates[5], uint32_t lut[5]) {
  for (int ro = 0; ro < 5; ++ro) {
    uint32_t word = (uint32_t)ro;
    for (int s = 1; s <= 4; ++s) {
      int target = rates[ro] / s;  // floor division of non-negative ints
      int best = 0, gap = abs(rates[0] - target);
      for (int i = 0; i < 5; ++i) {
        int g = abs(rates[i] - target);
        if (g < gap || (g == gap && rates[i] < rates[best])) { best = i; gap = g; }
      }
      word |= (uint32_t)best << (3 * s);
    }
    lut[ro] = word;
  }
}

// ---- a4: bandwidth-trace walk (bitrate_selection/simulators/network.py:22-35) --------------
struct TracePtr {   // read-only view of one trace row
  const double *p;
  MANSY_HD double operator()(int i) const {
#ifdef __CUDA_ARCH__
    return __ldg(p + i);
#else
    return p[i];
#endif
  }
};

// Returns the download time; advances (cur_idx, cur_time).  `ok` is cleared if the iteration cap
// is hit (a trace whose positive seconds are too small to ever finish the download).
template <typename ThrLoad>
MANSY_HD double trace_download(double size, ThrLoad thr_at, int trace_len, int &cur_idx, double &cur_time,
                               bool &ok) {
  const double start = cur_time;
  int it = 0;
  while (size > 0.0) {
    const double thr = thr_at(cur_idx);
    const double next_tick = floor(dadd(cur_time, 1.0));
    const double remain = dmul(dsub(next_tick, cur_time), thr);
    if (size >= remain) {
      cur_idx = (cur_idx + 1 == trace_len) ? 0 : cur_idx + 1;
      cur_time = next_tick;
      size = dsub(size, remain);
    } else {
      cur_time = dadd(cur_time, ddiv(size, thr));
      size = 0.0;
    }
    if (++it > (1 << 22)) { ok = false; break; }
  }
  return dsub(cur_time, start);
}

// ---- a5: playback buffer (bitrate_selection/simulators/buffer.py:8-15) ----------------------
MANSY_HD double buffer_push(double &buf, double chunk_length, double download_time) {
  if (download_time > buf) {
    const double rebuf = dsub(download_time, buf);
    buf = chunk_length;
    return rebuf;
  }
  buf = dadd(dsub(buf, download_time), chunk_length);
  return 0.0;
}

// ---- a8: QoE from the reduced sums (bitrate_selection/utils/qoe.py:22-34) --------------------
struct QoE {
  double qoe, q1, q2, q3;
};

MANSY_HD QoE qoe_from_sums(double vq, double sum_dev, double sum_m, double rebuffer, bool first_step,
                           double &prev_vq, double w0, double w1, double w2, double max_quality) {
  QoE r;
  const double intra = ddiv(ddiv(sum_dev, sum_m), max_quality);
  const double vqn = ddiv(vq, max_quality);
  const double inter = first_step ? 0.0 : fabs(dsub(vqn, prev_vq));
  prev_vq = vqn;
  r.q1 = vqn;
  r.q2 = rebuffer;
  r.q3 = dadd(intra, inter);
  r.qoe = dsub(dsub(dmul(w0, r.q1), dmul(w1, r.q2)), dmul(w2, r.q3));
  return r;
}

// The same QoE from the per-(chunk, action) parts: q1 = vq / max_quality and intra = sum_dev / sum_m / max_quality
// (qoe.py:23-28) are pure functions of the tables and the action; what a step adds is the rebuffer time and the
// inter-chunk term (qoe.py:29-33).  Same operations in the same order as qoe_from_sums.
MANSY_HD QoE qoe_from_parts(double q1, double intra, double rebuffer, bool first_step, double &prev_vq, double w0,
                            double w1, double w2) {
  QoE r;
  const double inter = first_step ? 0.0 : fabs(dsub(q1, prev_vq));
  prev_vq = q1;
  r.q1 = q1;
  r.q2 = rebuffer;
  r.q3 = dadd(intra, inter);
  r.qoe = dsub(dsub(dmul(w0, r.q1), dmul(w1, r.q2)), dmul(w2, r.q3));
  return r;
}

// ---- a13-a15: field of view -> tile mask (viewport_prediction/utils/common.py:37-58,83-127) ---
// blk(p) = p // b, minus one when p > 0 sits exactly on a tile boundary.
MANSY_HD int fov_block(int p, int b) {
  int k = p / b;
  if (p > 0 && p % b == 0) k -= 1;
  return k;
}

MANSY_HD uint32_t bit_range(int a, int b) {  // bits a..b inclusive (0 <= a, b <= 7), empty when a > b
  if (a > b) return 0u;
  return ((1u << (b + 1)) - 1u) & ~((1u << a) - 1u);
}

// One axis: tiles covered by [p-half, p+half] on a ring of `length` pixels cut into 8 tiles.
// Returns an 8-bit set; `valid` is cleared outside the reference's domain (centre outside the
// frame, or a FoV that would wrap on both sides: the reference raises there).
MANSY_HD uint32_t fov_axis_tiles(int p, int half, int length, int tile, bool &valid) {
  if (p < 0 || p > length) { valid = false; return 0u; }
  const int lo = p - half, hi = p + half;
  if (lo >= 0 && hi <= length) return bit_range(fov_block(lo, tile), fov_block(hi, tile));
  if (lo < 0 && hi <= length)   // wraps below 0: [0, hi] and [lo mod length, length]
    return bit_range(0, fov_block(hi, tile)) | bit_range(fov_block(lo + length, tile), fov_block(length, tile));
  if (lo >= 0 && hi > length)   // wraps above length: [0, hi mod length] and [lo, length]
    return bit_range(0, fov_block(hi - length, tile)) | bit_range(fov_block(lo, tile), fov_block(length, tile));
  valid = false;
  return 0u;
}

MANSY_HD uint64_t fov_tile_mask(int x, int y, int width, int height, int fov_w, int fov_h, bool &valid) {
  const uint32_t cols = fov_axis_tiles(x, fov_w / 2, width, width / 8, valid);
  const uint32_t rows = fov_axis_tiles(y, fov_h / 2, height, height / 8, valid);
  uint64_t m = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if ((rows >> r) & 1u) m |= (uint64_t)cols << (8 * r);
  return m;
}

// viewport_prediction/predict.py:40,43: int(value * video_width) with a float32 value; under the
// reference's pinned numpy the product is float64 (exact for a 24-bit x 12-bit product).
MANSY_HD int centre_to_pixel(float v, int length) { return (int)((double)v * (double)length); }

// ---- counter-based action stream (mansy_immersivevideostreaming_b200/synth.py:synthetic_actions) -
MANSY_HD int hashed_action(uint64_t seed, uint64_t env, uint64_t step, int n_actions) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ULL + env * 0xBF58476D1CE4E5B9ULL + step * 0x94D049BB133111EBULL;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27; z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (int)(z % (uint64_t)n_actions);
}

}  // namespace mansy
