/*
 * stitch_b200.h — C ABI of libstitchb200.so
 *
 * B200-native (sm_100a) kernels for the stitching-alignment hot path of
 * gargatik/Seamless-Through-Breaking-Rethinking-Image-Stitching-for-Optimal-Alignment:
 * FlowFormer's all-pairs cost volume / pyramid / lookup and the warp stage.
 *
 * The reference has NO FFI of its own (it is pure Python over ATen); the
 * "interface each entry point replaces" is therefore the Python function
 * named in each comment (paths relative to the reference root).  The host
 * side (package `stitch_b200`) mirrors those Python signatures and calls
 * these symbols through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch);
 *     the library never allocates or frees user-visible memory;
 *   - all tensors are dense, row-major ("contiguous"), fp32 unless noted;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on
 *     it, nothing synchronises, nothing touches the default stream;
 *   - return 0 on success, a negative SB_E* code on failure; the message is
 *     available (thread-local) through sb_last_error();
 *   - no CPU fallback: on a device that is not compute capability 10.x the
 *     compute entry points fail with SB_EARCH.
 */
#ifndef STITCH_B200_H
#define STITCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_EINVAL (-1)   /* bad shape / null pointer / alignment         */
#define SB_EARCH (-2)    /* device is not sm_100                         */
#define SB_ECUDA (-3)    /* a CUDA runtime / driver call failed          */
#define SB_EUNSUP (-4)   /* valid request outside the implemented range  */

#define SB_VERSION 100

typedef void* sb_stream_t;

/* ------------------------------------------------------------------ misc */
int sb_version(void);
const char* sb_last_error(void);
/* 0 if the current device is sm_100 (B200), SB_EARCH otherwise. */
int sb_device_check(void);
/* Number of kernels launched by this library since load / last reset
 * (bench.py's "gpu_launches"). */
long long sb_launch_count(void);
void sb_reset_launch_count(void);
/* Frees the library's own state (a 64-byte mapped debug word), resets sb_tune knobs and the launch counter.
 * Everything else (outputs, workspaces, sb_host_alloc blocks) is caller-owned.  The library stays usable. */
int sb_shutdown(void);
/* Pinned (page-locked, portable) host buffers for the host-resident data path; write_combined != 0
 * adds cudaHostAllocWriteCombined (CPU fills sequentially, device reads). NULL on failure. */
void* sb_host_alloc(size_t bytes, int write_combined);
int sb_host_free(void* p);
/* Tuning knobs for kernel experiments (tools/, bench sweeps). value 0 restores
 * the built-in default. Never changes results, only launch geometry (except the
 * knobs marked EXPERIMENT ONLY, which no product path sets). */
enum {
  SB_TUNE_LOOKUP_DEPTH = 0,        /* window groups in flight per warp: 1 or 2 (default) */
  SB_TUNE_LOOKUP_CTAS_PER_SM = 1,  /* resident CTAs per SM the persistent lookup grid is sized for */
  SB_TUNE_LOOKUP_SUPERBLOCK = 2,   /* queries per superblock: 8, 16 or 32 (default: by problem size) */
  SB_TUNE_LOOKUP_FETCH_ONLY = 3,   /* EXPERIMENT ONLY (changes results): windows are fetched, taps are not sampled */
  SB_TUNE_LOOKUP_L2_KEEP_EIGHTHS = 4, /* n/8 of the window fetches carry an L2 evict_last policy (default 3; 15 = no hint) */
  SB_TUNE_AGG_MBLK = 5,            /* attn @ v: 128-query blocks per CTA sharing each v stage: 1, 2 or 4 (default: by problem size) */
  SB_TUNE_CORR_2CTA = 6,           /* cost volume: 1 = one CTA per tile (default), 2 = CTA pairs (tcgen05 cta_group::2) */
  SB_TUNE_CORR_STORE_POLICY = 7,   /* cost volume: L2 policy of the output stores: 1 evict_first (default), 2 evict_last, 3 none */
  SB_TUNE_TPS_LOG = 8,             /* TPS basis log(): 0 = lg2.approx * ln2 (default), 1 = libdevice logf */
  SB_TUNE_LOOKUP_PDL = 9,          /* r = 4 lookup launched with programmatic stream serialization (prologue overlaps the previous kernel's tail): 0 off, 1 on */
  SB_TUNE_CORR_A_TMEM = 10,        /* cost volume: 1 (default) = the A block is copied to tensor memory once per unit and the MMAs read it from there, 2 = the MMAs read A from shared memory */
  SB_TUNE_LOOKUP_GENERIC = 11,     /* EXPERIMENT: 1 = r = 4 lookups take the generic window-staging kernel (LDG.128) instead of the TMA-box kernel */
  SB_TUNE_WARP_TILED = 12,         /* flow / homography warps: 1 = shared-memory-staged tiles where the shape allows (bit-identical, measured slower); default 0 = per-pixel gathers */
  SB_TUNE_CORR_TILES_PER_UNIT = 13, /* cost volume: target tiles per work unit (the A block is loaded once per unit): 4 (default), 8 or 16 */
  SB_TUNE_CORR_DYNAMIC = 14,       /* cost volume / attention logits: 1 (default) = work units handed out by the hardware scheduler (clusterlaunchcontrol.try_cancel, one CTA per unit in the grid), 2 = static round-robin over 148 persistent CTAs */
  SB_TUNE_COUNT = 16
};
int sb_tune(int key, int value);
/* Post-mortem word of the correlation kernel: 0, or 0xDEAD00tt when a bounded
 * mbarrier wait (tag tt) timed out and the kernel trapped instead of hanging. */
unsigned int sb_debug_word(void);

/* --------------------------------------------------------------- C1 / C2
 * Replaces MemoryEncoder.corr  (core/FlowFormer/PerCostFormer3/encoder.py:359-369)
 *   corr[b, i, j] = sum_d fmap1[b, d, i] * fmap2[b, d, j]     (heads = 1, no scale)
 * with bf16 operands, fp32 accumulation (tcgen05.mma), fp32 output.
 *
 * fmap1 [B, C, N1], fmap2 [B, C, N2] fp32 (NCHW with H*W flattened).
 * vol   [B, N1, N2] fp32  (== [B,1,H1,W1,H2,W2]).
 * lvl1/lvl2/lvl3: optional (may be NULL) avg-pool pyramid over the TARGET
 *   axes (H2, W2): lvl1 [B*N1, H2/2, W2/2], lvl2 [.., H2/4, W2/4],
 *   lvl3 [.., H2/8, W2/8]  == chained F.avg_pool2d(cost_maps, 2, stride=2)
 *   (C2; RAFT convention hinted at encoder.py:376 / common.py:245-248).
 *   Fused into the GEMM epilogue when W2 == 64 and H2 % 8 == 0, otherwise
 *   computed by the standalone pooling kernel after the volume.
 * workspace: caller-provided scratch of at least sb_corr_workspace_bytes()
 *   (bf16 token-major copies of both feature maps).
 * Limits: C <= 256 and C % 8 == 0 (SB_EUNSUP otherwise); N2 % 4 == 0.
 */
size_t sb_corr_workspace_bytes(int B, int C, int N1, int N2);
int sb_corr(const float* fmap1, const float* fmap2, float* vol,
            float* lvl1, float* lvl2, float* lvl3,
            void* workspace, size_t workspace_bytes,
            int B, int C, int H1, int W1, int H2, int W2, sb_stream_t stream);

/* The two halves of sb_corr, exposed so a caller can convert a feature map
 * once and reuse it for the forward and the backward volume.
 *   sb_feat_to_tokens_bf16: fmap [B, C, N] fp32 -> tok [B, N, Cpad] bf16,
 *     Cpad = C rounded up to 64, pad columns zero.
 *   sb_corr_tokens: volume (+ pyramid) from two token-major bf16 maps. */
int sb_feat_to_tokens_bf16(const float* fmap, void* tok, int B, int C, int N,
                           sb_stream_t stream);
int sb_corr_tokens(const void* tok1, const void* tok2, float* vol,
                   float* lvl1, float* lvl2, float* lvl3,
                   int B, int C, int H1, int W1, int H2, int W2, sb_stream_t stream);
/* Forward AND backward volume of Bp pairs in one launch (the two FlowFormer calls of flowHomoAdpater.py:158,178
 * share their inputs): tok_both [2*Bp, N, Cpad] bf16 = image-1 maps then image-2 maps; vol [2*Bp, N, N] (and the
 * levels) hold corr(img1_b, img2_b) for b < Bp and corr(img2_b, img1_b) for Bp + b.  H1 = H2 = H, W1 = W2 = W. */
int sb_corr_tokens_bidir(const void* tok_both, float* vol, float* lvl1, float* lvl2, float* lvl3,
                         int Bp, int C, int H, int W, sb_stream_t stream);
/* Opt-in (not the reference's dtype): the volume written as bf16 [B, N1, N2], H2*W2 % 8 == 0 — half
 * the HBM bytes, which moves the kernel from the write roofline towards the tensor pipe. */
int sb_corr_tokens_bf16out(const void* tok1, const void* tok2, void* vol_bf16,
                           int B, int C, int H1, int W1, int H2, int W2, sb_stream_t stream);
/* Same with an explicit row pitch of the volume (in floats, >= H2*W2, a multiple of 4): lets a
 * caller hold volumes whose token count is not a multiple of 4 in rows padded to the TMA's
 * 16-byte stride granularity (pad columns are never written; no pyramid in that case). */
int sb_corr_tokens_pitched(const void* tok1, const void* tok2, float* vol, long long vol_pitch,
                           float* lvl1, float* lvl2, float* lvl3,
                           int B, int C, int H1, int W1, int H2, int W2, sb_stream_t stream);

/* Standalone C2: out[p, y, x] = mean of the 2x2 block of in[p] (floor sizes,
 * like F.avg_pool2d(kernel 2, stride 2)).  in [P, H, W] -> out [P, H/2, W/2]. */
int sb_avg_pool2x2(const float* in, float* out, long long planes, int H, int W,
                   sb_stream_t stream);

/* ------------------------------------------------------------ C3 / C3p
 * Replaces MemoryDecoder.encode_flow_token (decoder.py:242-260) +
 * bilinear_sampler (core/utils/utils.py:62-76).
 *   cost_maps [B*H1*W1, H2, W2] fp32 (one map per query),
 *   coords    [B, 2, H1, W1] fp32, channel 0 = x, channel 1 = y,
 *   out       [B, H1, W1, (2r+1)^2] fp32 — the memory order the reference
 *             hands to its consumers (logical [B,(2r+1)^2,H1,W1]).
 *   out[q, i*(2r+1)+j] = bilinear(cost_maps[q], cx*scale + (i-r), cy*scale + (j-r))
 *   (slow window index steps x: RAFT's meshgrid(dy,dx) quirk), zeros padding,
 *   align_corners=True arithmetic of grid_sample.
 * coord_scale = 1 for the live single-level lookup; 1/2^l for pyramid level l
 * (C3p, common.py:245-248).  out_stride/out_offset (in floats) let several
 * levels interleave into one [.., L*(2r+1)^2] tensor; pass (2r+1)^2 and 0
 * for the plain case. r <= 7.
 */
int sb_corr_lookup(const float* cost_maps, const float* coords, float* out,
                   int B, int H1, int W1, int H2, int W2, int r,
                   float coord_scale, int out_stride, int out_offset,
                   sb_stream_t stream);

/* Generic bilinear_sampler (core/utils/utils.py:62-76):
 *   img [N, C, H, W], coords [N, Ho, Wo, 2] pixel (x,y) -> out [N, C, Ho, Wo]. */
int sb_bilinear_sampler(const float* img, const float* coords, float* out,
                        int N, int C, int H, int W, int Ho, int Wo,
                        sb_stream_t stream);

/* ------------------------------------------------------------------ W1
 * Replaces warp(x, flo) (core/warp_utils.py:71-80, bilinear mode):
 *   out[b,c,y,x] = bilinear(x[b,c], x + flo[b,0,y,x], y + flo[b,1,y,x]),
 *   zeros padding, align_corners=True, evaluated with the reference's exact
 *   normalise / un-normalise fp32 arithmetic.
 *   x [B,C,H,W], flo [B,2,H,W], out [B,C,H,W].
 * mul_mask (optional, [B,1,H,W]) is multiplied into every channel
 * (flowHomoAdpater.py:182,317 fused); NULL = plain warp.
 * overlap (optional, [B,H,W], needs C == 6): where(mean(out[:,3:6]) < 0.9, 1, 0)
 * of the UNMASKED warp (flowHomoAdpater.py:171-174 fused); NULL to skip. */
int sb_flow_warp(const float* x, const float* flo, const float* mul_mask, float* out,
                 float* overlap, int B, int C, int H, int W, sb_stream_t stream);
/* mode = 'nearest' of the same function (core/warp_utils.py:74-79 -> F.grid_sample(mode='nearest', zeros padding) — called
 * WITHOUT align_corners there, i.e. align_corners=False): the pixel at the half-to-even rounded sample position, 0 outside. */
int sb_flow_warp_nearest(const float* x, const float* flo, float* out, int B, int C, int H, int W,
                         sb_stream_t stream);

/* ------------------------------------------------------------------ W2
 * Replaces transformer(U, theta, out_size)
 * (core/udis_utils/torch_homo_transform.py:5-151).
 *   U [B,C,H,W]; theta [theta_batch,3,3] (theta_batch = 1 or B);
 *   xs [Wout], ys [Hout]: the torch.linspace(-1,1,n) tables (passed in, never
 *   recomputed — linspace is not reproducible arithmetically);
 *   out [B,C,Hout,Wout].
 *   idx_dbg: optional int32 [B,4,Hout,Wout] receiving the clamped integer
 *   grid indices (x0,x1,y0,y1) — the "integer grid indices" of the parity
 *   contract; NULL to skip.
 *   n_ones: the reference always warps cat(image, ones) (flowHomoAdpater.py:110-113,
 *   292,310,314); with n_ones > 0 the kernel behaves as if U had n_ones extra
 *   all-ones planes appended (out has C + n_ones channels) without reading them. */
int sb_homo_warp(const float* U, const float* theta, const float* xs, const float* ys,
                 float* out, int32_t* idx_dbg,
                 int B, int C, int n_ones, int H, int W, int Hout, int Wout, int theta_batch,
                 sb_stream_t stream);

/* ------------------------------------------------------------------ G1
 * Fused geometry of the adapter: tensor_DLT (core/udis_utils/torch_DLT.py:17-45)
 * + the normalised-coordinate products of core/flowHomoAdpater.py:96-113, one
 * launch, no host synchronisation (torch.inverse synchronises):
 *   H = DLT(src_p, dst_p) [B,3,3];  theta = L*H*R;  theta_inv = L*H^-1*R.
 * src_p, dst_p: DEVICE [B,4,2]; L_host, R_host: HOST pointers to 9 floats
 * (row-major 3x3, passed to the kernel by value); H / theta / theta_inv:
 * DEVICE [B,3,3], each optional (NULL to skip). */
int sb_dlt_theta(const float* src_p, const float* dst_p, const float* L_host,
                 const float* R_host, float* H, float* theta, float* theta_inv, int B,
                 sb_stream_t stream);

/* ------------------------------------------------------------------ W3
 * Replaces the dense part of transformer(U, source, target, out_size)
 * (core/udis_utils/torch_tps_transform.py:96-147 + sampler :18-94).
 *   T [B,2,pn+3] fp32: solved TPS coefficients (the (pn+3)^2 fp64 solve of
 *   :149-185 stays in torch — tiny, library LU);
 *   source [B,pn,2] control points in [-1,1]; xs/ys linspace tables. */
int sb_tps_warp(const float* U, const float* T, const float* source,
                const float* xs, const float* ys, float* out, int32_t* idx_dbg,
                int B, int C, int H, int W, int Hout, int Wout, int pn,
                sb_stream_t stream);

/* ------------------------------------------------------------------ W3k
 * Replaces warp_image_tps(image, kernel_centers, kernel_weights, affine_weights, align_corners)
 * (core/inference/tps_methods/kornia_tps.py:105-176; called from core/inference/tps_pipline.py:381):
 * dense kornia-style TPS evaluation on create_meshgrid(h, w) followed by
 * F.grid_sample(bilinear, zeros, align_corners) — fused, the [B, H*W, K] kernel matrix is never
 * materialised.
 *   centers, kweights [B,K,2]; affine [B,3,2]; xs [W], ys [H]: the meshgrid tables
 *   ((linspace(0, n-1, n) / (n-1) - 0.5) * 2, computed by the caller like the reference does);
 *   out [B,C,H,W]; coords_dbg: optional [B,H,W,2] warped sampling grid (NULL to skip). */
int sb_tps_kornia_warp(const float* image, const float* centers, const float* kweights,
                       const float* affine, const float* xs, const float* ys, float* out,
                       float* coords_dbg, int B, int C, int H, int W, int K, int align_corners,
                       sb_stream_t stream);
/* F.grid_sample(img [N,C,H,W], grid [N,Ho,Wo,2], mode='bilinear', padding_mode='zeros',
 * align_corners) -> out [N,C,Ho,Wo]  (the sampler of kornia_tps.py:172). */
int sb_grid_sample(const float* img, const float* grid, float* out, int N, int C, int H, int W,
                   int Ho, int Wo, int align_corners, sb_stream_t stream);

/* Attention.forward in two passes over the same tcgen05 contraction (core/FlowFormer/PerCostFormer3/gma.py:54-76):
 * attn[b, i, :] = tf32(softmax_j(tok_q[b, i, :] . tok_k[b, j, :])) without a round trip of the fp32 logits
 * through HBM: pass 1 keeps the running (max, sum of exp) of every query row, pass 2 recomputes the
 * contraction and writes normalised probabilities.  tok_q / tok_k: bf16 token-major maps from
 * sb_feat_to_tokens_bf16 ([B, N, Cpad], scale folded into q); attn [B, Nq, Nk] fp32 (Nk % 4 == 0);
 * stats: workspace of B * Nq * 2 floats (8-byte aligned). */
int sb_attn_softmax_tokens(const void* tok_q, const void* tok_k, float* attn, float* stats, int B, int C,
                           int Nq, int Nk, sb_stream_t stream);
/* ------------------------------------------------------------------ N1 ("next" row 1)
 * GMA attention / aggregation (core/FlowFormer/PerCostFormer3/gma.py:54-76, :102-115).
 *   sb_softmax_rows: in-place softmax over rows of a dense fp32 matrix (the `sim.softmax(dim=-1)`
 *     of gma.py:73 applied to the q.k^T volume produced by sb_corr); to_tf32 != 0 rounds the
 *     probabilities to TF32 (round-to-nearest) so sb_attn_aggregate's tensor-core read is unbiased.
 *   sb_attn_aggregate: out[bh, n, i] = (residual[bh, n, i] +) gamma * sum_j attn[bh, i, j] * v[bh, n, j]
 *     attn [BH, Nq, Nk], v [BH, d, Nk] (the 1x1-conv layout), out / residual [BH, d, Nq]; d = 128;
 *     residual and gamma (DEVICE pointer to one float) are optional (NULL). */
int sb_softmax_rows(float* x, long long rows, int n, long long row_stride, int to_tf32,
                    sb_stream_t stream);
int sb_attn_aggregate(const float* attn, const float* v, const float* residual, const float* gamma,
                      float* out, int BH, int Nq, int Nk, int d, sb_stream_t stream);
/* Opt-in half-traffic variant (not the reference's dtype): sb_softmax_rows_bf16 writes the
 * probabilities as a dense bf16 matrix [rows, n] (x is left untouched), sb_attn_aggregate_bf16
 * consumes bf16 attn [BH, Nq, Nk] and bf16 v [BH, d, Nk] (kind::f16 MMA); Nk % 8 == 0. */
int sb_softmax_rows_bf16(const float* x, void* out_bf16, long long rows, int n, long long row_stride,
                         sb_stream_t stream);
int sb_attn_aggregate_bf16(const void* attn_bf16, const void* v_bf16, const float* residual,
                           const float* gamma, float* out, int BH, int Nq, int Nk, int d,
                           sb_stream_t stream);

/* D[bh] = A[bh] . B[bh]^T on the TF32 tensor cores (the kernel behind sb_attn_aggregate, row-major
 * output): A [BH, M, K], B [BH, N, K] fp32 K-major, D [BH, M, N] fp32; K % 4 == 0. */
int sb_gemm_nt_tf32(const float* A, const float* B, float* D, int BH, int M, int N, int K,
                    sb_stream_t stream);

/* ------------------------------------------------------------------ N3 ("next" row 3)
 * Replaces UDIS2Network.CCL(feature_1, feature_2)  (core/UDIS2/Homography/network.py:147-199):
 * L2-normalise over channels, 3x3-patch all-pairs correlation, softmax(10 x) over the patches,
 * expected displacement. feature_1/2 [B,C,H,W] -> flow [B,2,H,W] (channel 0 = w, 1 = h).
 * workspace: caller-provided, 256-byte aligned, >= sb_ccl_workspace_bytes(). C % 4 == 0; H*W unbounded when H*W % 4 == 0 and
 * W <= 256 (the staged kernel), else H*W <= 4096 (the row-streaming fallback). */
size_t sb_ccl_workspace_bytes(int B, int C, int H, int W);
int sb_ccl(const float* feature_1, const float* feature_2, float* flow, void* workspace,
           size_t workspace_bytes, int B, int C, int H, int W, float softmax_scale, sb_stream_t stream);

/* ------------------------------------------------------------------ N2 ("next" row 2)
 * Replaces MemoryDecoder.upsample_flow(flow, mask)
 * (core/FlowFormer/PerCostFormer3/decoder.py:214-225, called every GRU iteration at :331):
 * convex 8x upsampling, softmax over the 9 taps of mask [N, 576, H, W] (viewed [N,1,9,8,8,H,W])
 * applied to the 3x3 zero-padded neighbourhood of 8 * flow [N, 2, H, W] -> out [N, 2, 8H, 8W]. */
int sb_upsample_flow(const float* flow, const float* mask, float* out, int N, int H, int W,
                     sb_stream_t stream);

/* ------------------------------------------------------------------ N4 ("next" row 4)
 * Replaces the convolution stack of PatchEmbed.forward
 * (core/FlowFormer/PerCostFormer3/encoder.py:36-43 built, :68-73 run; called per cost map at :263):
 *   Conv2d(1,16,6,stride 2,pad 2) -> ReLU -> Conv2d(16,32,6,2,2) -> ReLU -> Conv2d(32,64,6,2,2)
 * cost_maps [n_maps, 1, 64, 64] fp32 -> out [n_maps, 64, 8, 8] fp32.  bf16 operands, fp32 accumulation
 * (tcgen05, CTA pairs); activations between the layers are rounded to bf16.
 *   sb_patch_embed_pack: once per set of weights — w1 [16,1,6,6], w2 [32,16,6,6], w3 [64,32,6,6] fp32 ->
 *     `pack` (sb_patch_embed_pack_bytes() bytes, 16-byte aligned), the kernel's shared-memory images.
 *   bias: b1 | b2 | b3 = 16 + 32 + 64 fp32, contiguous.
 * Only 64 x 64 maps (512 x 512 images, the shipped configuration) are supported: SB_EUNSUP otherwise. */
size_t sb_patch_embed_pack_bytes(void);
int sb_patch_embed_pack(const float* w1, const float* w2, const float* w3, void* pack, sb_stream_t stream);
int sb_patch_embed_proj(const float* cost_maps, const void* pack, const float* bias, float* out,
                        long long n_maps, int H, int W, sb_stream_t stream);

/* ------------------------------------------------------------------ W4
 * Replaces compute_range_map(flow) (core/warp_utils.py:114-175): forward
 * bilinear splat count.  Deterministic: weights are accumulated as 2^-32
 * fixed-point integers (order independent), then converted to fp32.
 *   flow [B,2,H,W]; accum: caller scratch, B*H*W uint64; range_map [B,1,H,W].
 * mode 0: range_map = raw count
 * mode 1: range_map = clamp(count,0,1)            (compute_occlusion 'wang', occlusion_are_zeros=True, :213-220)
 * mode 2: range_map = 1 - clamp(count,0,1)        (occlusion_are_zeros=False)
 * mode 3: range_map = clamp(count,0,1) >= 0.5     (fused caller threshold, flowHomoAdpater.py:181) */
int sb_range_map(const float* flow, unsigned long long* accum, float* range_map,
                 int B, int H, int W, int mode, sb_stream_t stream);

/* ------------------------------------------------------------------ W5
 * Replaces preprocess_occlusion_mask(mask, kernel_size=(kh,kw))
 * (core/flowHomoAdpater.py:18-35): binarise >= 0.5, erode, dilate with a
 * kh x kw box (zero padded), binarise.  mask/out [P,H,W] fp32 (P = B*1).
 * Also serves the cv2 11x11 open of core/inference/tps_pipline.py:143-147
 * with border_is_zero = 0 (cv2 treats outside as "no constraint"). */
int sb_morph_open(const float* mask, float* out, int P, int H, int W,
                  int kh, int kw, int border_is_zero, sb_stream_t stream);

/* ------------------------------------------------------------------ W6
 * Replaces the compositing of test_out_forward
 * (core/flowHomoAdpater.py:317,333-360), one fused pass:
 *   fw   = final_warp (already * flow_mask) * occ
 *   out2 = H2[:, :3]*(1-m2)*(1-m1) + fw[:, :3]*m2      m2 = fw[:,3:6]
 *   m2'  = H2[:,3:6]*(1-m2)*(1-m1) + m2*m2
 *   blend = clip((o1*m1 + out2*m2')/(m1+m2'), 0, 255) -> uint8
 *   mask1 = clip(mean_c m1,0,1) x3 ; mask2 = clip(mean_c m2',0,1) x3
 * homo1 [B,6,H,W] (img1 | mask), homo2 [B,6,H,W], fw_in [B,6,H,W] (warp of
 * homo2 by the residual flow, already multiplied by the flow mask),
 * occ [B,1,H,W] or NULL (the use_fb_consistency_mask=False branch :347-351).
 * Outputs: final_warp [B,6,H,W] (fw), output2 [B,3,H,W], mask1, mask2
 * [B,3,H,W] fp32, blend [B,3,H,W] uint8. */
int sb_composite_test_out(const float* homo1, const float* homo2, const float* fw_in,
                          const float* occ, float* final_warp, float* output2,
                          float* mask1, float* mask2, uint8_t* blend,
                          int B, int H, int W, sb_stream_t stream);

/* ------------------------------------------------------------------ W7
 * Replaces the arithmetic of build_model (core/UDIS2/Composition/network.py:8-20);
 * `net_out` [B,1,H,W] is the UNet output (not ours). All others [B,3,H,W]. */
int sb_build_model(const float* warp1, const float* warp2, const float* mask1,
                   const float* mask2, const float* net_out,
                   float* learned_mask1, float* learned_mask2, float* stitched,
                   int B, int H, int W, sb_stream_t stream);

/* ------------------------------------------------------------------ W8
 * Replaces the TPS-stage mix + average blend of
 * core/inference/tps_pipline.py:150-170 (elementwise part):
 *   fm   = (mean_c(final_warp >= 3) >= .5) ; inv1 = (mean_c(1-mask1) >= .5)
 *   tfw  = final_warp*fm + tps_warp*(1-fm)*inv1
 *   tfm  = fm + (1-fm)*tps_mask*inv1
 *   out2 = tfw*tfm ; blend = clip((o1*m1 + out2*tfm)/(m1+tfm),0,255) -> u8
 * final_warp, tps_warp (already * tps_mask), output1, mask1 [B,3,H,W];
 * tps_mask [B,1,H,W].  Outputs: output2 [B,3,H,W], mask2 [B,1,H,W], blend u8. */
int sb_tps_mix_blend(const float* final_warp, const float* tps_warp, const float* tps_mask,
                     const float* output1, const float* mask1,
                     float* output2, float* mask2, uint8_t* blend,
                     int B, int H, int W, sb_stream_t stream);

/* ------------------------------------------------- train_eval fused tail
 * Replaces flowHomoAdpater.py:171-174 (+:182): overlap = mean_c(mask) < 0.9,
 * from final_warp_output [B,6,H,W] -> overlap [B,H,W]. */
int sb_overlap_mask(const float* final_warp, float* overlap, int B, int H, int W,
                    sb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STITCH_B200_H */
