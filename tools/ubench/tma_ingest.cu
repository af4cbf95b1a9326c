// Micro-benchmark: per-SM operand ingest rate from L2 on sm_100a.
//   mode 0: TMA 2D boxes [rows x 32 floats] SWIZZLE_128B into a ring of `stages` buffers, one CTA per SM.
//   mode 1: LDG.128 -> STS.128 copy by 128 threads of the same bytes.
//   mode 2: cp.async (LDGSTS.128) by 128 threads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest tma_ingest.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(160, 1) ingest(const __grid_constant__ CUtensorMap map, const float *src, int mode, int stages, int box_rows,
                                                 int n_boxes, int total_cols_boxes, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t box_bytes = box_rows * 128;
  const uint32_t bars = base + stages * box_bytes * (mode == 4 ? 2 : 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    if (threadIdx.x == 0) {
      for (int it = 0; it < n_boxes + stages; ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(bars + 8 * s, ((it - stages) / stages) & 1);     // previous fill of this stage landed
        if (it < n_boxes) {
          mbar_expect_tx(bars + 8 * s, box_bytes);
          const int b = (it + blockIdx.x * 7) % total_cols_boxes;
          tma_load_2d(base + s * box_bytes, &map, b * 32, 0, bars + 8 * s);
        }
      }
    }
  } else if (mode == 3) {
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < 2) {
      for (int it = w; it < n_boxes + stages; it += 2) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(bars + 8 * s, ((it - stages) / stages) & 1);
        if (it < n_boxes) {
          mbar_expect_tx(bars + 8 * s, box_bytes);
          const int b = (it + blockIdx.x * 7) % total_cols_boxes;
          tma_load_2d(base + s * box_bytes, &map, b * 32, 0, bars + 8 * s);
        }
      }
    }
  } else if (mode == 4) {     // one thread, two TMA ops of half a box each per stage (like an L1 job)
    if (threadIdx.x == 0) {
      for (int it = 0; it < n_boxes + stages; ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(bars + 8 * s, ((it - stages) / stages) & 1);
        if (it < n_boxes) {
          mbar_expect_tx(bars + 8 * s, box_bytes * 2);
          const int b = (it + blockIdx.x * 7) % (total_cols_boxes - 1);
          tma_load_2d(base + s * box_bytes * 2, &map, b * 32, 0, bars + 8 * s);
          tma_load_2d(base + s * box_bytes * 2 + box_bytes, &map, b * 32 + 32, 0, bars + 8 * s);
        }
      }
    }
  } else {
    // 128 copy threads: each box = box_rows rows x 8 chunks of 16 B
    const int t = threadIdx.x;
    if (t < 128) {
      const int row_stride_f = total_cols_boxes * 32;
      float4 acc = make_float4(0, 0, 0, 0);
      for (int it = 0; it < n_boxes; ++it) {
        const int s = it % stages;
        const int b = (it + blockIdx.x * 7) % total_cols_boxes;
        const uint32_t dst = base + s * box_bytes;
        for (int idx = t; idx < box_rows * 8; idx += 128) {
          const int r = idx >> 3, c = idx & 7;
          const float *g = src + (size_t)r * row_stride_f + b * 32 + c * 4;
          const uint32_t d = dst + r * 128 + ((c ^ (r & 7)) << 4);
          if (mode == 1) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(g));
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(d), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          } else {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
          }
        }
        if (mode == 2) {
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 3;" ::: "memory");
        }
      }
      if (mode == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (acc.x == 1234.5f) cycles[1] = 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = clock64() - t0;
}

int main(int argc, char **argv) {
  const int rows = 256, col_boxes = 40;            // [256 rows][1280 floats] = 1.31 MB, L2 resident
  float *d; cudaMalloc(&d, (size_t)rows * col_boxes * 32 * 4); cudaMemset(d, 0, (size_t)rows * col_boxes * 32 * 4);
  long long *cyc; cudaMalloc(&cyc, 16);
  void *fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Fn fn = (Fn)fnp;
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int grids[3] = {1, 32, 148};
  const int ng = 2;
  for (int box_rows : {64, 128, 256}) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)col_boxes * 32, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)col_boxes * 32 * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int mode : {0, 3, 4})
      for (int stages : {2, 4, 6})
        for (int g = 0; g < ng; ++g) {
          const int box_bytes = box_rows * 128;
          if (stages * box_bytes * (mode == 4 ? 2 : 1) > 192 * 1024) continue;
          const int n_boxes = (4 << 20) / box_bytes;     // 4 MB per CTA
          for (int rep = 0; rep < 2; ++rep) {
            ingest<<<grids[g], 160, stages * box_bytes * (mode == 4 ? 2 : 1) + 2048>>>(map, d, mode, stages, box_rows, n_boxes, col_boxes, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
          }
          long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
          printf("mode %d box_rows %3d stages %d grid %3d : %7.1f B/clk/SM  (%lld cycles for 4 MB)\n", mode, box_rows, stages, grids[g],
                 (double)n_boxes * box_bytes * (mode == 4 ? 2 : 1) / c, c);
        }
  }
  return 0;
}
