// composite.cu — W6 / W7 / W8: fused overlap-mask compositing and blending.
// Pure streaming kernels (every input plane read once, every output written
// once); one pixel per thread with all loads issued before the arithmetic.
// Arithmetic follows the reference expression by expression (one fp32
// rounding per torch op, evaluated left to right).
#include "common.cuh"

namespace sb {

__device__ __forceinline__ float clipf(float v, float lo, float hi) {
  // torch.clip propagates NaN
  return (v != v) ? v : fminf(fmaxf(v, lo), hi);
}
__device__ __forceinline__ uint8_t to_u8(float v) {
  // .to(torch.uint8) of the CPU oracle: truncate; NaN -> 0
  if (v != v) return 0;
  return (uint8_t)(int)v;
}
__device__ __forceinline__ float mean3(float a, float b, float c) {
  return fdiv(fadd(fadd(a, b), c), 3.0f);
}

// ---------------------------------------------------------------- W6
// core/flowHomoAdpater.py:337-360 (with occ) and :347-351 (without).
__global__ void __launch_bounds__(256)
composite_test_out_kernel(const float* __restrict__ homo1, const float* __restrict__ homo2,
                          const float* __restrict__ fw_in, const float* __restrict__ occ,
                          float* __restrict__ final_warp, float* __restrict__ output2,
                          float* __restrict__ mask1, float* __restrict__ mask2,
                          uint8_t* __restrict__ blend, long long plane, long long total) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const long long b = p / plane, rem = p - b * plane;
    const float* h1 = homo1 + b * 6 * plane + rem;
    const float* h2 = homo2 + b * 6 * plane + rem;
    const float* fw = fw_in + b * 6 * plane + rem;
    float a1[6], a2[6], f[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      a1[c] = ldg_stream(h1 + c * plane);
      a2[c] = ldg_stream(h2 + c * plane);
      f[c] = ldg_stream(fw + c * plane);
    }
    const bool has_occ = occ != nullptr;
    const float o = has_occ ? ldg_stream(occ + b * plane + rem) : 1.0f;
    if (has_occ) {
#pragma unroll
      for (int c = 0; c < 6; ++c) f[c] = fmul(f[c], o);     // :337 final_warp_output * occlusion_mask
    }
    float o2[3], m2n[3], m1v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float m1c = a1[3 + c], m2c = f[3 + c];
      const float one_m2 = fsub(1.0f, m2c);
      m1v[c] = m1c;
      if (has_occ) {
        const float non_ov = fsub(1.0f, m1c);               // :341
        o2[c] = fadd(fmul(fmul(a2[c], one_m2), non_ov), fmul(f[c], m2c));          // :343
        m2n[c] = fadd(fmul(fmul(a2[3 + c], one_m2), non_ov), fmul(m2c, m2c));      // :344
      } else {
        o2[c] = fadd(fmul(a2[c], one_m2), fmul(f[c], m2c));                         // :348
        m2n[c] = fadd(fmul(a2[3 + c], one_m2), fmul(m2c, m2c));                     // :349
      }
    }
    float* fwo = final_warp + b * 6 * plane + rem;
#pragma unroll
    for (int c = 0; c < 6; ++c) stg_stream(fwo + c * plane, f[c]);
    const float mm1 = clipf(mean3(m1v[0], m1v[1], m1v[2]), 0.0f, 1.0f);            // :359
    const float mm2 = clipf(mean3(m2n[0], m2n[1], m2n[2]), 0.0f, 1.0f);            // :360
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o3 = (b * 3 + c) * plane + rem;
      stg_stream(output2 + o3, o2[c]);
      stg_stream(mask1 + o3, mm1);
      stg_stream(mask2 + o3, mm2);
      // :355-356  (o1*m1 + o2*m2) / (m1 + m2) -> clip -> uint8
      const float num = fadd(fmul(a1[c], m1v[c]), fmul(o2[c], m2n[c]));
      const float den = fadd(m1v[c], m2n[c]);
      blend[o3] = to_u8(clipf(fdiv(num, den), 0.0f, 255.0f));
    }
  }
}

// ---------------------------------------------------------------- W7
// core/UDIS2/Composition/network.py:12-14
__global__ void __launch_bounds__(256)
build_model_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                   const float* __restrict__ m1, const float* __restrict__ m2,
                   const float* __restrict__ net_out, float* __restrict__ lm1,
                   float* __restrict__ lm2, float* __restrict__ st, long long plane,
                   long long total /* B*3*plane */) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long bc = i / plane, rem = i - bc * plane;
    const long long b = bc / 3;
    const float a = ldg_stream(m1 + i), c = ldg_stream(m2 + i);
    const float o = __ldg(net_out + b * plane + rem);
    const float x1 = ldg_stream(w1 + i), x2 = ldg_stream(w2 + i);
    const float mm = fmul(a, c);
    const float l1 = fadd(fsub(a, mm), fmul(mm, o));
    const float l2 = fadd(fsub(c, mm), fmul(mm, fsub(1.0f, o)));
    const float s = fsub(fadd(fmul(fadd(x1, 1.0f), l1), fmul(fadd(x2, 1.0f), l2)), 1.0f);
    stg_stream(lm1 + i, l1);
    stg_stream(lm2 + i, l2);
    stg_stream(st + i, s);
  }
}

// ---------------------------------------------------------------- W8
// core/inference/tps_pipline.py:150-170
__global__ void __launch_bounds__(256)
tps_mix_blend_kernel(const float* __restrict__ final_warp, const float* __restrict__ tps_warp,
                     const float* __restrict__ tps_mask, const float* __restrict__ output1,
                     const float* __restrict__ mask1, float* __restrict__ output2,
                     float* __restrict__ mask2, uint8_t* __restrict__ blend, long long plane,
                     long long total) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const long long b = p / plane, rem = p - b * plane;
    float fw[3], tp[3], o1[3], m1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o3 = (b * 3 + c) * plane + rem;
      fw[c] = ldg_stream(final_warp + o3);
      tp[c] = ldg_stream(tps_warp + o3);
      o1[c] = ldg_stream(output1 + o3);
      m1[c] = ldg_stream(mask1 + o3);
    }
    const float tm = ldg_stream(tps_mask + b * plane + rem);
    const float fm_mean = mean3(fw[0] >= 3.0f ? 1.0f : 0.0f, fw[1] >= 3.0f ? 1.0f : 0.0f,
                                fw[2] >= 3.0f ? 1.0f : 0.0f);                        // :151
    const float fm = fm_mean >= 0.5f ? 1.0f : 0.0f;                                  // :152
    const float inv_mean = mean3(fsub(1.0f, m1[0]), fsub(1.0f, m1[1]), fsub(1.0f, m1[2]));  // :154
    const float inv1 = inv_mean >= 0.5f ? 1.0f : 0.0f;                               // :155
    const float one_fm = fsub(1.0f, fm);
    const float tfm = fadd(fm, fmul(fmul(one_fm, tm), inv1));                        // :157
    stg_stream(mask2 + b * plane + rem, tfm);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o3 = (b * 3 + c) * plane + rem;
      const float tfw = fadd(fmul(fw[c], fm), fmul(fmul(tp[c], one_fm), inv1));      // :156
      const float o2 = fmul(tfw, tfm);                                               // :162
      stg_stream(output2 + o3, o2);
      const float num = fadd(fmul(o1[c], m1[c]), fmul(o2, tfm));                     // :168
      const float den = fadd(m1[c], tfm);
      blend[o3] = to_u8(clipf(fdiv(num, den), 0.0f, 255.0f));                        // :169
    }
  }
}

// -------------------------------------------------- train_eval overlap mask
// core/flowHomoAdpater.py:171-174
__global__ void __launch_bounds__(256)
overlap_mask_kernel(const float* __restrict__ final_warp, float* __restrict__ overlap,
                    long long plane, long long total) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const long long b = p / plane, rem = p - b * plane;
    const float* m = final_warp + (b * 6 + 3) * plane + rem;
    const float mean = mean3(ldg_stream(m), ldg_stream(m + plane), ldg_stream(m + 2 * plane));
    stg_stream(overlap + p, mean < 0.9f ? 1.0f : 0.0f);
  }
}

static inline int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long max_blocks = (long long)kNumSMs * 8 * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  return (int)blocks;
}

}  // namespace sb

extern "C" int sb_composite_test_out(const float* homo1, const float* homo2, const float* fw_in,
                                     const float* occ, float* final_warp, float* output2,
                                     float* mask1, float* mask2, uint8_t* blend, int B, int H,
                                     int W, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();

  SB_REQUIRE(B >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_composite_test_out: bad size");
  const long long plane = (long long)H * W, total = plane * B;
  if (total == 0) return SB_OK;
  SB_REQUIRE(homo1 && homo2 && fw_in && final_warp && output2 && mask1 && mask2 && blend,
             SB_EINVAL, "sb_composite_test_out: null pointer");
  composite_test_out_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
      homo1, homo2, fw_in, occ, final_warp, output2, mask1, mask2, blend, plane, total);
  SB_LAUNCH_CHECK("composite_test_out_kernel");
  return SB_OK;
}

extern "C" int sb_build_model(const float* warp1, const float* warp2, const float* mask1,
                              const float* mask2, const float* net_out, float* learned_mask1,
                              float* learned_mask2, float* stitched, int B, int H, int W,
                              sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();

  SB_REQUIRE(B >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_build_model: bad size");
  const long long plane = (long long)H * W, total = plane * B * 3;
  if (total == 0) return SB_OK;
  SB_REQUIRE(warp1 && warp2 && mask1 && mask2 && net_out && learned_mask1 && learned_mask2 && stitched,
             SB_EINVAL, "sb_build_model: null pointer");
  build_model_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
      warp1, warp2, mask1, mask2, net_out, learned_mask1, learned_mask2, stitched, plane, total);
  SB_LAUNCH_CHECK("build_model_kernel");
  return SB_OK;
}

extern "C" int sb_tps_mix_blend(const float* final_warp, const float* tps_warp,
                                const float* tps_mask, const float* output1, const float* mask1,
                                float* output2, float* mask2, uint8_t* blend, int B, int H, int W,
                                sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();

  SB_REQUIRE(B >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_tps_mix_blend: bad size");
  const long long plane = (long long)H * W, total = plane * B;
  if (total == 0) return SB_OK;
  SB_REQUIRE(final_warp && tps_warp && tps_mask && output1 && mask1 && output2 && mask2 && blend,
             SB_EINVAL, "sb_tps_mix_blend: null pointer");
  tps_mix_blend_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
      final_warp, tps_warp, tps_mask, output1, mask1, output2, mask2, blend, plane, total);
  SB_LAUNCH_CHECK("tps_mix_blend_kernel");
  return SB_OK;
}

extern "C" int sb_overlap_mask(const float* final_warp, float* overlap, int B, int H, int W,
                               sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();

  SB_REQUIRE(B >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_overlap_mask: bad size");
  const long long plane = (long long)H * W, total = plane * B;
  if (total == 0) return SB_OK;
  SB_REQUIRE(final_warp && overlap, SB_EINVAL, "sb_overlap_mask: null pointer");
  overlap_mask_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(final_warp, overlap, plane, total);
  SB_LAUNCH_CHECK("overlap_mask_kernel");
  return SB_OK;
}
