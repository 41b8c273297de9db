// Write-bandwidth probe for the cost-volume store pattern: [B, N1, N2] fp32, N1 = N2 = 4096, B = 16.
// Each persistent CTA walks (b, 128-row block, 4 consecutive 128-column tiles) like corr_umma_kernel and
// only stores. Variants:
//   0: 4 warps, TMA box 32 cols x 32 rows (4 KB), 2 staging buffers per warp   (= corr_umma_kernel today)
//   1: 4 warps, TMA box 32 cols x 128 rows (16 KB), one warp per column slice
//   2: st.global.v4, a warp writes one 512-byte tile row per instruction
//   3: 1 warp, cp.async.bulk (non-tensor) 512 B per tile row
//   4: as 0 but the unit is 1 tile (tiles of a row block visited by different CTAs)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o store_probe store_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int N = 4096, B = 16;

__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(K) : "memory"); }
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int VARIANT>
__global__ void __launch_bounds__(128, 1)
store_kernel(const __grid_constant__ CUtensorMap map32, const __grid_constant__ CUtensorMap map128, float* vol);

// pyramid variants: vol + lvl1 [B*N, 32, 32] + lvl2 [B*N,16,16] + lvl3 [B*N,8,8]
//   6: today's kernel: per tile lvl1 TMA 32x32 box (128 B per query), lvl2 64 B per query every 2 tiles (st.global),
//      lvl3 32 B per query every 4 tiles
//   7: per unit (4 tiles): lvl1 512 B per query in one TMA store (box 128 x 32 rows), lvl2 128 B per query,
//      lvl3 32 B per query
template <int VARIANT>
__global__ void __launch_bounds__(128, 1)
pyr_kernel(const __grid_constant__ CUtensorMap map32, const __grid_constant__ CUtensorMap map_l1a,
           const __grid_constant__ CUtensorMap map_l1b, float* l2, float* l3) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = N / 128, MB = N / 128, NG = NT / 4;
  const long long n_units = (long long)B * MB * NG;
  int sbuf = 0;
  const float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
  for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int ng = (int)(u % NG);
    const long long r1 = u / NG;
    const int mb = (int)(r1 % MB), b = (int)(r1 / MB);
    const long long q = (long long)b * N + mb * 128 + warp * 32 + lane;
    for (int tt = 0; tt < 4; ++tt) {
      const int t = ng * 4 + tt;
      for (int sl = 0; sl < 4; ++sl) {
        if (lane == 0) wait_read<1>();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&map32, sbase + (warp * 2 + sbuf) * 4096, t * 128 + sl * 32, mb * 128 + warp * 32, b);
          commit();
        }
        sbuf ^= 1;
      }
      if (VARIANT == 6) {
        if (lane == 0) wait_read<1>();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&map_l1a, sbase + (warp * 2 + sbuf) * 4096, t * 32, mb * 128 + warp * 32, b);
          commit();
        }
        sbuf ^= 1;
        if (tt & 1) {
          float* o = l2 + (q * 16 + (t >> 1)) * 16;
          for (int c = 0; c < 4; ++c) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
        if (tt == 3) {
          float* o = l3 + (q * 8 + (t >> 2)) * 8;
          for (int c = 0; c < 2; ++c) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
    }
    if (VARIANT == 7) {
      if (lane == 0) {
        wait_read<1>();
        tma_store_3d(&map_l1b, sbase + 32768 + warp * 16384, ng * 128, mb * 128 + warp * 32, b);
        commit();
      }
      float* o = l2 + (q * 16 + ng * 2) * 16;
      for (int c = 0; c < 8; ++c) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      float* o3 = l3 + (q * 8 + ng) * 8;
      for (int c = 0; c < 2; ++c) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o3 + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
  wait_all();
}

template <int VARIANT>
__global__ void __launch_bounds__(128, 1)
store_kernel(const __grid_constant__ CUtensorMap map32, const __grid_constant__ CUtensorMap map128, float* vol) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_unit = (VARIANT == 4) ? 1 : 4;
  const int NT = N / 128, MB = N / 128, NG = NT / tiles_per_unit;
  const long long n_units = (long long)B * MB * NG;
  int sbuf = 0;
  for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int ng = (int)(u % NG);
    const long long r1 = u / NG;
    const int mb = (int)(r1 % MB), b = (int)(r1 / MB);
    for (int tt = 0; tt < tiles_per_unit; ++tt) {
      const int t = ng * tiles_per_unit + tt;
      if (VARIANT == 0 || VARIANT == 4) {
        for (int sl = 0; sl < 4; ++sl) {
          if (lane == 0) wait_read<1>();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map32, sbase + (warp * 2 + sbuf) * 4096, t * 128 + sl * 32, mb * 128 + warp * 32, b);
            commit();
          }
          sbuf ^= 1;
        }
      } else if (VARIANT == 8 || VARIANT == 9) {
        constexpr int NB = VARIANT == 8 ? 4 : 8;     // staging buffers per warp (16 / 32 KB in flight per warp)
        for (int sl = 0; sl < 4; ++sl) {
          if (lane == 0) wait_read<NB - 1>();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map32, sbase + (warp * NB + sbuf) * 4096, t * 128 + sl * 32, mb * 128 + warp * 32, b);
            commit();
          }
          sbuf = (sbuf + 1) % NB;
        }
      } else if (VARIANT == 1) {
        if (lane == 0) wait_read<1>();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&map128, sbase + (warp * 2 + sbuf) * 16384, t * 128 + warp * 32, mb * 128, b);
          commit();
        }
        sbuf ^= 1;
      } else if (VARIANT == 2) {
        float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
        for (int r = warp; r < 128; r += 4) {
          float* p = vol + ((long long)b * N + mb * 128 + r) * N + t * 128 + lane * 4;
          asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      } else if (VARIANT == 3) {
        if (warp == 0) {
          // 128 rows of 512 B, lane l issues rows l, l+32, ...
          wait_read<0>();
          for (int r = lane; r < 128; r += 32) {
            float* p = vol + ((long long)b * N + mb * 128 + r) * N + t * 128;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(p), "r"(sbase + r * 512) : "memory");
          }
          commit();
        }
      }
    }
  }
  if (VARIANT != 2) wait_all();
}

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  float* vol; cudaMalloc(&vol, (size_t)B * N * N * 4);
  CUtensorMap m32, m128;
  cuuint64_t dims[3] = {N, N, B}; cuuint64_t str[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
  cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t box32[3] = {32, 32, 1}, box128[3] = {32, 128, 1};
  enc(&m32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vol, dims, str, box32, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&m128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vol, dims, str, box128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = 160 * 1024;
  cudaEvent_t e0_, e1_; cudaEventCreate(&e0_); cudaEventCreate(&e1_);
  cudaFuncSetAttribute(store_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(store_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(store_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(store_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(store_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(store_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(store_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int v = 8; v <= 9; ++v) {
    for (int i = 0; i < 13; ++i) {
      if (i == 3) cudaEventRecord(e0_);
      if (v == 8) store_kernel<8><<<148, 128, smem>>>(m32, m128, vol);
      else store_kernel<9><<<148, 128, smem>>>(m32, m128, vol);
    }
    cudaEventRecord(e1_);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0_, e1_); ms /= 10;
    printf("variant %d (%d staging buffers per warp): %7.1f us  %6.0f GB/s\n", v, v == 8 ? 4 : 8, ms * 1e3, (double)B * N * N * 4 / ms / 1e6);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  {
    float *l1, *l2, *l3;
    cudaMalloc(&l1, (size_t)B * N * 1024 * 4); cudaMalloc(&l2, (size_t)B * N * 256 * 4); cudaMalloc(&l3, (size_t)B * N * 64 * 4);
    CUtensorMap ml1a, ml1b;
    cuuint64_t d1[3] = {1024, N, B}; cuuint64_t s1[2] = {1024 * 4, (cuuint64_t)N * 1024 * 4};
    cuuint32_t bxa[3] = {32, 32, 1}, bxb[3] = {128, 32, 1};
    enc(&ml1a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, l1, d1, s1, bxa, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r = enc(&ml1b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, l1, d1, s1, bxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("l1b encode failed %d\n", (int)r);
    cudaFuncSetAttribute(pyr_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(pyr_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int v = 6; v <= 7; ++v) {
      for (int i = 0; i < 13; ++i) {
        if (i == 3) cudaEventRecord(e0);
        if (v == 6) pyr_kernel<6><<<148, 128, smem>>>(m32, ml1a, ml1b, l2, l3);
        else pyr_kernel<7><<<148, 128, smem>>>(m32, ml1a, ml1b, l2, l3);
      }
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pyr variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
      printf("pyramid variant %d: %7.1f us  %6.0f GB/s\n", v, ms * 1e3, (double)B * N * (N + 1024 + 256 + 64) * 4 / ms / 1e6);
    }
  }
  for (int v = 0; v < 6; ++v) {
    for (int grid : {148, 296}) {
      if (v == 5) {  // cudaMemset reference
        if (grid != 148) continue;
        for (int i = 0; i < 3; ++i) cudaMemsetAsync(vol, 0, (size_t)B * N * N * 4);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) cudaMemsetAsync(vol, 0, (size_t)B * N * N * 4);
        cudaEventRecord(e1);
      } else {
        auto launch = [&]() {
          switch (v) {
            case 0: store_kernel<0><<<grid, 128, (grid == 148 ? smem : 100 * 1024)>>>(m32, m128, vol); break;
            case 1: store_kernel<1><<<grid, 128, (grid == 148 ? smem : 100 * 1024)>>>(m32, m128, vol); break;
            case 2: store_kernel<2><<<grid, 128, 1024>>>(m32, m128, vol); break;
            case 3: store_kernel<3><<<grid, 128, (grid == 148 ? smem : 100 * 1024)>>>(m32, m128, vol); break;
            case 4: store_kernel<4><<<grid, 128, (grid == 148 ? smem : 100 * 1024)>>>(m32, m128, vol); break;
          }
        };
        if (v == 1 && grid == 296) continue;   // 8 x 16 KB staging does not fit twice
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) launch();
        cudaEventRecord(e1);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d grid %d: %s\n", v, grid, cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
      printf("variant %d grid %3d: %7.1f us  %6.0f GB/s\n", v, grid, ms * 1e3, (double)B * N * N * 4 / ms / 1e6);
    }
  }
  return 0;
}
