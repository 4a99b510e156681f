/*
 * xpoint_b200.h -- C ABI of the B200-native (sm_100a) XPoint inference hot path.
 *
 * This is the drop-in boundary: every entry point takes plain device pointers, sizes,
 * element strides, dtype enums and a CUDA stream; there are no torch / C++ types in any
 * signature.  The library never allocates device memory, never synchronises the stream
 * and keeps no global state except a thread-local error string.  All functions return
 * XP_OK (0) or a negative xp_status; xp_last_error() gives the text.
 *
 * Each entry point names the reference interface (file:line in canyagmur/XPoint) it
 * replaces; INTEGRATION.md shows the reference-side ctypes binding for each.
 */
#ifndef XPOINT_B200_H_
#define XPOINT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XP_ABI_VERSION 5

#if defined(__GNUC__)
#define XP_API __attribute__((visibility("default")))
#else
#define XP_API
#endif

typedef enum {
    XP_OK = 0,
    XP_ERR_INVALID_ARG = -1,   /* shape / stride / dtype / null-pointer violation (reference: TORCH_CHECK) */
    XP_ERR_UNSUPPORTED = -2,   /* valid in the reference, not implemented here (e.g. backward)            */
    XP_ERR_CUDA = -3,          /* a CUDA runtime / launch error; text in xp_last_error()                   */
    XP_ERR_WORKSPACE = -4      /* caller-provided workspace too small                                      */
} xp_status;

typedef enum { XP_F32 = 0, XP_F16 = 1, XP_BF16 = 2 } xp_dtype;

typedef void* xp_stream_t; /* cudaStream_t */

/* -- library ------------------------------------------------------------------------- */
XP_API int xp_abi_version(void);
XP_API const char* xp_last_error(void);              /* thread-local, valid until the next failing call */
/* 0 when the current CUDA device can run this library (compute capability 10.x). */
XP_API int xp_check_device(void);

/* -- a1/a2: selective scan forward ----------------------------------------------------
 * Replaces selective_scan_cuda_oflex.fwd  (kernels/selective_scan/csrc/selective_scan/
 * cusoflex/selective_scan_oflex.cpp:143-231, kernel selective_scan_fwd_kernel_oflex.cuh:67-212)
 * behind selective_scan_fn (csms6s.py:112-126), plus the mamba_ssm-style extras exercised by
 * the kernel tests (z gate, last state: test_selective_scan.py:168-234).
 *
 *   delta' = softplus(delta + delta_bias)            (softplus optional; threshold 20)
 *   h_l    = exp(delta'_l * A) * h_{l-1} + delta'_l * B_l * u_l ,   h_{-1} = 0
 *   out_l  = sum_n C_l[n] * h_l[n] + D * u_l ;   out_l *= silu(z_l)   (z optional)
 *
 * u, out, z : (batch, dim, seqlen)            delta : (batch, delta_dim, seqlen), dim % delta_dim == 0
 * A : (dim, dstate) fp32 contiguous           B, C  : (batch, groups, dstate, seqlen), dim % groups == 0
 * D : (dim) fp32 or NULL                      delta_bias : (delta_dim) fp32 or NULL
 * last_state : (batch, dim, dstate) fp32 contiguous or NULL.
 * u/delta/B/C/z share in_dtype; out has out_dtype (XP_F32, or == in_dtype).  All strides are in
 * ELEMENTS; the seqlen stride of every tensor must be 1 (same rule as selective_scan_oflex.cpp:170-176).
 * dstate <= 256.  State and accumulation are always fp32.
 *
 * Fused CrossScan addressing (ABI 2; the three trailing fields, zero-initialised = the contract above):
 * with u_group_div > 0 the u row of (batch b, group g, row dg inside the group) is
 *     u + b*u_batch_stride + (g / u_group_div)*u_group_stride + dg*u_dim_stride ,
 * so several scan directions read ONE copy of the activations (CrossScan without the 4x copy:
 * csm_triton.py:22-29 routes l0 / l1 / L-1-l0 / L-1-l1 -- directions 0|2 share x, 1|3 share x^T).
 * Bit g of reverse_group_mask makes group g (g < 64) walk the sequence backwards through memory: every
 * per-token tensor of that group (u, delta, B, C, z) is read at token L-1-l at scan step l and out is written
 * there, i.e. its y stays in natural (unflipped) memory order (CrossMerge's flips, csm_triton.py:56-62).
 *
 * Fused dt_proj (ABI 3; SURVEY 8f row f1 -- VMamba.py:605-615 computes delta = dt_projs_weight x dts_r and hands the
 * materialised (batch, dim, seqlen) tensor to the scan).  With dt_rank > 0 the scan takes the low-rank factors instead
 * and the (batch, dim, seqlen) delta never exists in HBM: `delta` points at dts_r (batch, groups, dt_rank, seqlen) in
 * in_dtype with element strides delta_batch_stride / dt_group_stride / delta_dim_stride (= rank-row stride), dt_weight is
 * (dim, dt_rank) contiguous in in_dtype, and
 *     delta_l[d] = round_to_in_dtype( sum_r dt_weight[d, r] * dts_r[b, g(d), r, l] )      (fp32 accumulation)
 * before delta_bias / softplus -- the value the reference's dt_proj GEMM produces under autocast.  delta_dim must equal
 * dim.  dt_rank <= 64; 16-bit inputs with dt_rank <= 16 and dstate <= 2 run the rank-R product on the tensor cores inside
 * the scan kernel (mma.sync, one MMA along the rank), everything else takes the shape-generic kernel.
 */
typedef struct {
    const void* u;
    const void* delta;
    const float* A;
    const void* B;
    const void* C;
    const float* D;          /* nullable */
    const void* z;           /* nullable */
    const float* delta_bias; /* nullable */
    void* out;
    float* last_state;       /* nullable */
    int64_t batch, dim, delta_dim, groups, dstate, seqlen;
    int64_t u_batch_stride, u_dim_stride;
    int64_t delta_batch_stride, delta_dim_stride;
    int64_t B_batch_stride, B_group_stride, B_state_stride;
    int64_t C_batch_stride, C_group_stride, C_state_stride;
    int64_t z_batch_stride, z_dim_stride;
    int64_t out_batch_stride, out_dim_stride;
    int32_t in_dtype;        /* xp_dtype */
    int32_t out_dtype;       /* xp_dtype */
    int32_t delta_softplus;  /* bool */
    int32_t force_generic;   /* debug/testing: 1 = always take the shape-generic kernel */
    /* fused CrossScan addressing, see above (all zero = the plain selective_scan_fn contract) */
    int64_t u_group_stride;       /* elements */
    int64_t u_group_div;          /* 0 = classic addressing */
    uint64_t reverse_group_mask;  /* bit g = group g runs backwards through memory */
    /* fused dt_proj, see above (all zero = delta is materialised) */
    const void* dt_weight;        /* (dim, dt_rank) in_dtype, contiguous */
    int64_t dt_rank;
    int64_t dt_group_stride;      /* elements */
} xp_scan_args;

XP_API int xp_selective_scan_fwd(const xp_scan_args* args, xp_stream_t stream);

/* The reference extension also exports bwd (selective_scan_oflex.cpp:233-355).  This library is
 * inference-only: the symbol exists so a binding can resolve it, and always returns
 * XP_ERR_UNSUPPORTED. */
XP_API int xp_selective_scan_bwd(void);

/* -- a3/a4: CrossScan / CrossMerge ----------------------------------------------------
 * Replace cross_scan_fn / cross_merge_fn and the CrossScanF / CrossMergeF (+Triton) forward
 * passes (csm_triton.py:22-179, 182-273, 278-517).
 *   scans 0: l0=h*W+w, l1=w*H+h, l2=L-1-l0, l3=L-1-l1 ; 1: four copies ; 2: two forward + two flipped.
 * x : channel-first (B,C,H,W) or channel-last (B,H,W,C); with one_by_one (B,4,C,H,W) / (B,H,W,4,C).
 * xs: channel-first (B,4,C,L) or channel-last (B,L,4,C).
 * merge: ys as xs above (viewed (B,4,C,H,W) / (B,H,W,4,C)); y channel-first (B,C,L) or channel-last
 * (B,L,C); with one_by_one y is (B,4,C,L) / (B,L,4,C) and nothing is summed.
 * Merge sums in fp32 with the torch path's association (y0 + y2') + (y1' + y3') and writes dtype.
 */
XP_API int xp_cross_scan(const void* x, void* xs, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype,
                  int32_t in_channel_first, int32_t out_channel_first, int32_t one_by_one, int32_t scans,
                  xp_stream_t stream);
XP_API int xp_cross_merge(const void* ys, void* y, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype,
                   int32_t in_channel_first, int32_t out_channel_first, int32_t one_by_one, int32_t scans,
                   xp_stream_t stream);

/* -- a4+a5 tail: CrossMerge (+) out_norm LayerNorm (+) z gate, one pass ------------------
 * Replaces cross_merge_fn -> transpose -> out_norm [-> y * z]   (VMamba.py:632-646 and :364-372).
 * ys : (B, 4, C, L) scan outputs in scan order, fp32 or 16-bit.   gamma/beta : (C) fp32.
 * zact : (B, L, C) already-activated gate (SiLU(z)) in out dtype, or NULL.
 * out  : (B, L, C) i.e. (B,H,W,C) channel-last, out_dtype.   eps: LayerNorm epsilon (1e-5).
 * workspace: xp_merge_norm_gate_workspace_bytes(B, C, H, W) bytes (the merged fp32 activations).
 */
XP_API int64_t xp_merge_norm_gate_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W);
XP_API int xp_merge_norm_gate(const void* ys, const float* gamma, const float* beta, const void* zact, void* out,
                       int64_t B, int64_t C, int64_t H, int64_t W, int32_t ys_dtype, int32_t out_dtype, float eps,
                       void* workspace, int64_t workspace_bytes, xp_stream_t stream);

/* -- a3+a4+a5 fused: the copy-free SS2D core around xp_selective_scan_fwd ------------------
 * Replaces cross_scan_fn + cross_merge_fn + out_norm of SS2D.forward_corev2 / forwardv0 (VMamba.py:603,632-646,
 * :364-372) without materialising the four scan orders: the scan itself routes the directions
 * (xp_scan_args.u_group_div = 2, reverse_group_mask = 0b1010) and these entry points provide the two layouts it
 * reads and the one pass that consumes its four outputs.  Direction order everywhere below is
 * [row-major forward (k=0), row-major backward (k=2), column-major forward (k=1), column-major backward (k=3)].
 *
 * xp_ss2d_pack:        x (B, D, H, W) -> xx (B, 2, D, L) = [x ; x^T] (x^T: token (h,w) at w*H+h), any dtype.
 * xp_ss2d_dwconv_pack: the same preceded by SS2D's depth-wise 3x3 convolution (padding 1) and SiLU
 *                      (VMamba.py:651-655): in is the CHANNEL-LAST in_proj output, channel c of token (h,w) at
 *                      in[((b*H+h)*W+w)*in_token_stride + c], c < D; weight (D, 3, 3) fp32, bias (D) fp32 or NULL.
 * xp_ss2d_merge_norm:  ys (B, 4, D, L) fp32, every plane in natural (unflipped) memory order, planes 0|1 row-major
 *                      and 2|3 column-major -> out (B, H, W, D) = LayerNorm_D(ys0 + ys1 + (ys2 + ys3)^T) * gamma
 *                      + beta [* zact], zact (B, H, W, D) in out_dtype or NULL.  H % 4 == W % 4 == 0, D <= 3072.
 */
XP_API int xp_ss2d_pack(const void* x, void* xx, int64_t B, int64_t D, int64_t H, int64_t W, int32_t dtype,
                        xp_stream_t stream);
XP_API int xp_ss2d_dwconv_pack(const void* in, const float* weight, const float* bias, void* xx, int64_t B, int64_t D,
                               int64_t H, int64_t W, int64_t in_token_stride, int32_t dtype, int32_t silu,
                               xp_stream_t stream);
 /* xp_ss2d_dt_proj: delta (B, G, D, L) = weight (G, D, R) fp32 x dts_r (B, G, R, L): the dt_proj of SS2D (VMamba.py:607-608,
 * :325) as a store-bound outer product.  R = dt_rank <= 8 for any dtype; 9..16 for 16-bit inputs (weights rounded to the
 * input dtype as under autocast, mixed-precision FMA with fp32 accumulation).  dts_r is a strided view of the x_proj
 * output (element strides given); delta is contiguous, in the dtype of dts_r. */
XP_API int xp_ss2d_dt_proj(const void* dts_r, const float* weight, void* delta, int64_t B, int64_t G, int64_t D, int64_t R,
                           int64_t L, int64_t x_batch_stride, int64_t x_group_stride, int64_t x_rank_stride, int32_t dtype,
                           xp_stream_t stream);
XP_API int xp_ss2d_merge_norm(const float* ys, const float* gamma, const float* beta, const void* zact, void* out,
                              int64_t B, int64_t D, int64_t H, int64_t W, int32_t out_dtype, float eps,
                              xp_stream_t stream);

/* -- a1+a3+a4 fused (ABI 4): the SS2D core with CrossMerge inside the scan ------------------
 * Replaces cross_scan_fn -> selective_scan_fn -> cross_merge_fn of SS2D.forward_corev2 (VMamba.py:603-632; CrossMerge
 * csm_triton.py:56-85) for d_state 1 | 2: the four directional scans of a channel accumulate into shared-memory token
 * planes and the MERGED y (B, D, H, W) fp32 = ys0 + flip(ys2) + (ys1 + flip(ys3))^T is written once -- no per-direction
 * y planes in HBM.  Inputs are those of the copy-free path above, in the fused direction order:
 *   xx (B, 2, D, L) = [x ; x^T] (xp_ss2d_dwconv_pack / xp_ss2d_pack), delta (B, 4, D, L) (xp_ss2d_dt_proj; softplus and
 *   delta_bias are applied here), B / C (B, 4, N, L) as strided views (element strides; unit stride along L),
 *   A (4*D, N), D (4*D) or NULL, delta_bias (4*D) or NULL, all fp32.
 * xp_ss2d_core_channels returns how many channels of one image a CTA holds for this shape, or 0 when a token plane does
 * not fit in shared memory (then use xp_selective_scan_fwd + xp_ss2d_merge_norm).
 * xp_ss2d_plane_norm: y (B, D, H, W) fp32 (the output above) -> out (B, H, W, D) = LayerNorm_D(y) * gamma + beta [* zact]:
 * out_norm (+ gate) of VMamba.py:641-646 / :369-372 reading 4 bytes per element. */
typedef struct xp_ss2d_core_args {
    const void* xx; const void* delta; const void* B; const void* C;
    const float* A; const float* D; const float* delta_bias;
    void* out;
    int64_t batch, d_inner, dstate, H, W;
    int64_t B_batch_stride, B_group_stride, B_state_stride, C_batch_stride, C_group_stride, C_state_stride;
    int32_t in_dtype, delta_softplus;
} xp_ss2d_core_args;
XP_API int32_t xp_ss2d_core_channels(int64_t d_inner, int64_t dstate, int64_t H, int64_t W, int32_t in_dtype);
XP_API int xp_ss2d_core(const xp_ss2d_core_args* args, xp_stream_t stream);
XP_API int xp_ss2d_plane_norm(const float* y, const float* gamma, const float* beta, const void* zact, void* out,
                              int64_t B, int64_t D, int64_t H, int64_t W, int32_t out_dtype, float eps,
                              xp_stream_t stream);

/* -- f2 (ABI 4): Linear + bias + residual add + LayerNorm in one tcgen05 GEMM ---------------
 * Replaces  pend = A W^T + b ; x = x + pend ; n = LayerNorm(x)  at the end of a VSSBlock branch (SS2D.out_proj VMamba.py:664 /
 * Mlp.fc2 :110-128 followed by the block's residual add and the next norm, :1222-1234):
 *   x_new (M, N) fp32 = residual + A (M, K) W (N, K)^T + bias   (x_new may be NULL)
 *   y     (M, N) 16-bit = LayerNorm_N(x_new) * gamma + beta
 * A, W, y in `dtype` (fp16 | bf16), residual / x_new / bias / gamma / beta fp32.  N must be 96, 192 or 384 (the whole row sits
 * in one TMEM accumulator tile), K % 8 == 0, all tensors contiguous and 16-byte aligned. */
XP_API int xp_linear_res_ln(const void* A, const void* W, const float* bias, const float* residual, const float* gamma,
                            const float* beta, float* x_new, void* y, int64_t M, int64_t N, int64_t K, int32_t dtype, float eps,
                            xp_stream_t stream);

/* -- f2 (ABI 5): the whole Mlp branch of a VSSBlock in one tcgen05 kernel ---------------
 * Replaces  h = GELU(n W1^T + b1) ; pend = h W2^T + b2 ; x = x + pend ; n' = LayerNorm(x)  (Mlp.forward VMamba.py:110-128 with
 * the block's residual add and the next norm, :1229-1234); the (M, 4C) hidden activation stays in shared memory.
 *   A (M, C) 16-bit, W1 (4C, C), W2 (C, 4C) 16-bit, b1 (4C) / b2 (C) fp32 or NULL, residual (M, C) fp32
 *   x_new (M, C) fp32 = residual + fc2(GELU(fc1(A)))   (may be NULL);  y (M, C) 16-bit = LayerNorm_C(x_new) * gamma + beta,
 *   or with gamma == NULL  y = x_new rounded to `dtype` (the stage's last block: a strided convolution follows, VMamba.py:906-919)
 * C must be 96 or 192; exact (erf) GELU; the hidden activation is rounded to `dtype` between the two GEMMs, as the
 * reference's autocast does.  All tensors contiguous and 16-byte aligned. */
XP_API int xp_mlp_res_ln(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, const float* residual,
                         const float* gamma, const float* beta, float* x_new, void* y, int64_t M, int64_t C, int32_t dtype,
                         float eps, xp_stream_t stream);

/* -- f3 / f4 ("next" rows): the evaluation driver's per-sample geometry, batched on the device --------------
 * Replaces warp_keypoints + filter_points (xpoint/utils/homographies.py:479-495,511-526: cv2.perspectiveTransform in float64
 * on (x, y), numpy astype(int)) and the per-sample loops of compute_repeatability_for_sample / compute_descriptor_for_sample
 * (xpoint/utils/benchmark_evaluation.py:396-467,650-690).  keypoints (B, k, 2) int32 (y, x) with counts (B) (rows >= count
 * ignored), H (B, 3, 3) float64 row-major.
 * xp_warp_keypoints: out_float (B, k, 2) float64 (y', x') and / or out_int (B, k, 2) int32 (truncated toward zero) and / or
 *   inside (B, k) uint8 = filter_points on the int result (filter_on_float = 0) or on the float result (1); any may be NULL.
 * xp_repeatability_counts: for set A already warped to B's frame (out_int + inside above): counts (B, n_thr) = points of A
 *   inside the image whose nearest keypoint of set B is within thresholds[t] (float64, <= 8 of them), n_inside (B).
 * xp_match_score_counts: queries warped by the ground-truth homography in float64 (out_float + inside with filter_on_float):
 *   n_correct (B, n_thr) matches (q, match_idx[q]) with || float32(warped_q - kp_t) || <= thr, n_gt (B, n_thr) queries with at
 *   least one train keypoint within thr, n_possible (B) warped queries inside the image, n_matches (B). */
XP_API int xp_warp_keypoints(const int32_t* keypoints, const int32_t* count, const double* H, int64_t B, int64_t k,
                             int64_t height, int64_t width, double* out_float, int32_t* out_int, uint8_t* inside,
                             int32_t filter_on_float, xp_stream_t stream);
XP_API int xp_repeatability_counts(const int32_t* warped_a, const uint8_t* inside_a, const int32_t* count_a,
                                   const int32_t* keypoints_b, const int32_t* count_b, int64_t B, int64_t k,
                                   const double* thresholds, int64_t n_thresholds, int32_t* counts, int32_t* n_inside,
                                   xp_stream_t stream);
XP_API int xp_match_score_counts(const double* warped_q, const uint8_t* inside_q, const int32_t* count_q,
                                 const int32_t* keypoints_t, const int32_t* count_t, const int32_t* match_idx, int64_t B,
                                 int64_t k, const double* thresholds, int64_t n_thresholds, int32_t* n_correct, int32_t* n_gt,
                                 int32_t* n_possible, int32_t* n_matches, xp_stream_t stream);

/* -- f2 (first "next" row): channel-last LayerNorm ---------------------------------------
 * Replaces the nn.LayerNorm calls around the SS2D block (VSSBlock.norm / norm2, patch-embed and downsample norms:
 * VMamba.py:1222-1234, :1405-1440).  x (rows, C) in in_dtype -> y (rows, C) in out_dtype; gamma/beta (C) fp32;
 * fp32 statistics, biased variance, eps inside the square root (torch semantics).  C <= 1536.
 */
XP_API int xp_layer_norm(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int64_t C,
                         int32_t in_dtype, int32_t out_dtype, float eps, xp_stream_t stream);

/* -- f2: residual add + bias + LayerNorm in one pass; patch-embed stem -----------------------
 * xp_add_layer_norm replaces `x = x + branch; n = LayerNorm(x)` of VSSBlock (VMamba.py:1222-1234) and the
 * `conv bias -> permute -> LayerNorm` tails of patch-embed / downsample (VMamba.py:1405-1440):
 *     s = x [+ res] [+ pre_bias] ;  sum_out = s (optional) ;  y = LayerNorm_C(s) * gamma + beta (optional)
 * x, res, y, sum_out: (rows, C) channel-last, each with its own dtype; pre_bias / gamma / beta: (C) fp32.  C <= 1536.
 * xp_patch_embed_stem replaces cat(x,x,x) -> Conv2d(3 -> C1, k3, s2, p1) -> permute -> LayerNorm(C1) -> permute -> GELU
 * (VMamba.py:1405-1413, :1509-1510): img (B, Cin, H, W) fp32 with Cin 1 (weights pre-summed over the replicated
 * channels by the caller) or 3, weight (C1, Cin, 3, 3) fp32 -> out (B, ceil(H/2), ceil(W/2), C1) channel-last.
 */
XP_API int xp_add_layer_norm(const void* x, const void* res, const float* pre_bias, const float* gamma, const float* beta,
                             void* y, void* sum_out, int64_t rows, int64_t C, int32_t x_dtype, int32_t res_dtype,
                             int32_t y_dtype, int32_t sum_dtype, float eps, xp_stream_t stream);
XP_API int xp_patch_embed_stem(const float* img, const float* weight, const float* bias, const float* gamma,
                               const float* beta, void* out, int64_t B, int64_t Cin, int64_t H, int64_t W, int64_t C1,
                               float eps, int32_t out_dtype, int32_t gelu, xp_stream_t stream);

/* -- f2: Linear (+ bias) (+ exact GELU) on tcgen05 tensor cores --------------------------------
 * Replaces `fc1 -> GELU` of the VSSBlock MLP (VMamba.py:110-128, :1229-1233): out (M, N) = act(A (M, K) W (N, K)^T + bias),
 * A / W / out in `dtype` (XP_F16 | XP_BF16), bias (N) fp32 or NULL, fp32 accumulation; gelu != 0 applies the exact (erf)
 * GELU to the accumulator so the hidden activation is written once.  K % 8 == 0, N % 32 == 0, row-major contiguous.
 */
XP_API int xp_linear_act(const void* A, const void* W, const float* bias, void* out, int64_t M, int64_t N, int64_t K,
                         int32_t dtype, int32_t gelu, xp_stream_t stream);

/* -- a6: detector post ----------------------------------------------------------------
 * Replaces Softmax2d -> [:, :-1] -> PixelShuffle(r)   (XPoint.py:356-357).
 * logits (B, r*r+1, Hc, Wc) in `dtype` -> prob (B, 1, r*Hc, r*Wc) fp32.  r <= 8.
 */
XP_API int xp_detector_post(const void* logits, float* prob, int64_t B, int64_t Hc, int64_t Wc, int32_t r,
                     int32_t dtype, xp_stream_t stream);

/* -- a7: descriptor head L2 normalisation ---------------------------------------------
 * Replaces F.normalize(x, p=2, dim=1)   (XPoint.py:365-366).   x (B, C, HW) in `dtype` -> fp32.
 * out_cf (B, C, HW) and/or out_cl (B, HW, C) (channel-last, the layout xp_sample_descriptors likes);
 * either may be NULL.
 */
XP_API int xp_l2_normalize(const void* x, float* out_cf, float* out_cl, int64_t B, int64_t C, int64_t HW, int32_t dtype,
                    xp_stream_t stream);

/* Channel-last forms (the drop-in model runs the heads' 1x1 convolutions as GEMMs on (cells, C) rows; these read the GEMM
 * outputs where they lie):
 *   xp_detector_post_cl: logits (B*Hc*Wc, ld) rows, the first r*r+1 entries of a row are the cell's logits (ld >= r*r+1,
 *                        e.g. 72 for a 16-byte aligned 65-wide GEMM) -> prob (B, 1, r*Hc, r*Wc) fp32.
 *   xp_l2_normalize_cl:  x (B, HW, C) rows -> out_cf (B, C, HW) and/or out_cl (B, HW, C) fp32, unit L2 norm per row.
 *   xp_encoder_tail:     the VSSM output path feeding the heads, one pass: s = x [+ pend] on the last stage's channel-last
 *                        (B, H, W, C_out*bs*bs) residual stream (x fp32, pend any dtype or NULL) -> depth_to_space(bs)
 *                        (VMamba.py:1500-1505) -> enc_out (B, C_out, H*bs, W*bs) fp32 channel-first (the model's
 *                        `encoder_output`, XPoint.py:309; nullable) and `padded` (B, H*bs+2, W*bs+2, C_out) channel-last in
 *                        pad_dtype = ReflectionPad2d(1) of it (XPoint.py:112,125; nullable).  C_out in {8, 16, 32, 48, 64}. */
XP_API int xp_detector_post_cl(const void* logits, float* prob, int64_t B, int64_t Hc, int64_t Wc, int32_t r, int64_t ld,
                               int32_t dtype, xp_stream_t stream);
XP_API int xp_l2_normalize_cl(const void* x, float* out_cf, float* out_cl, int64_t B, int64_t C, int64_t HW, int32_t dtype,
                              xp_stream_t stream);
XP_API int xp_encoder_tail(const void* x, const void* pend, float* enc_out, void* padded, int64_t B, int64_t H, int64_t W,
                           int64_t C_out, int64_t bs, int32_t x_dtype, int32_t pend_dtype, int32_t pad_dtype,
                           xp_stream_t stream);

/* -- a8/a9: greedy box NMS + top-k + keypoint compaction ------------------------------
 * Replaces utils.box_nms (xpoint/utils/utils.py:148-192, i.e. torchvision.ops.nms/batched_nms on
 * size x size boxes centred on every pixel above min_prob) and torch.nonzero(prob_nms > thr)
 * (xpoint/utils/evaluation.py:281-282).  Bit-exact with sequential greedy NMS (ties: lower flat index
 * wins).  prob (B, H, W) fp32 -> prob_nms (B, H, W) fp32 (zeros + kept scores; nullable).
 * keypoints (B, kp_capacity, 2) int32 (y, x) in raster order and kp_count (B) int32 are optional
 * (nullable); pixels counted are those with prob_nms > kp_threshold; kp_count holds the true count even
 * if it exceeds kp_capacity (only the first kp_capacity are written).
 * workspace: xp_nms_workspace_bytes(B, H, W) bytes of device memory, 16-byte aligned (state + scan-cursor byte planes).
 */
XP_API int64_t xp_nms_workspace_bytes(int64_t B, int64_t H, int64_t W);
XP_API int xp_box_nms(const float* prob, float* prob_nms, int64_t B, int64_t H, int64_t W, float size, float min_prob,
               float iou, int64_t keep_top_k, float kp_threshold, int32_t* keypoints, int32_t* kp_count,
               int64_t kp_capacity, void* workspace, int64_t workspace_bytes, xp_stream_t stream);

/* -- a10: bilinear descriptor sampling + L2 normalisation -----------------------------
 * Replaces utils.interpolate_descriptors (xpoint/utils/utils.py:229-238: grid_sample bilinear,
 * align_corners=True, zeros padding, then F.normalize).
 * keypoints (B, kp_stride, 2) int32 (y, x); kp_count (B) int32 or NULL (then every image has kp_stride).
 * desc: channel-first (B, C, Hc, Wc) or channel-last (B, Hc, Wc, C) fp32.   out (B, kp_stride, C) fp32;
 * rows >= kp_count[b] are zero-filled.  H, W: full-resolution image size.
 */
XP_API int xp_sample_descriptors(const int32_t* keypoints, const int32_t* kp_count, int64_t B, int64_t kp_stride,
                          const float* desc, int32_t channel_last, int64_t C, int64_t Hc, int64_t Wc, int64_t H,
                          int64_t W, float* out, xp_stream_t stream);

/* -- a11: mutual nearest-neighbour matching -------------------------------------------
 * Replaces get_matches(d1, d2, 'bfmatcher', crossCheck=True) (xpoint/utils/matching.py:4-36,
 * cv2.BFMatcher(NORM_L2, crossCheck=True).match) and NNMatcher.match (matching.py:38-75).
 * d1 (P, n1_stride, C), d2 (P, n2_stride, C) fp32, C % 32 == 0, C <= 512; n1/n2 (P) int32 valid rows per pair
 * (NULL = all).  Similarity GEMM on tcgen05 tensor cores (3xTF32 split, fp32-level accuracy), row / column
 * arg-min of the squared L2 distance fused into the epilogue, ties -> lowest index.
 * Outputs: nn12 (P, n1_stride) int32 = argmin_j, nn21 (P, n2_stride) int32 = argmin_i,
 *          match_idx (P, n1_stride) int32 = nn12[i] if mutual else -1, match_dist (P, n1_stride) fp32 =
 *          exact fp32 L2 distance of mutual pairs (recomputed on CUDA cores), match_count (P) int32.
 * Any output pointer may be NULL.  workspace: xp_match_workspace_bytes(P, n1_stride, n2_stride, C).
 * use_tensor_cores = 0 selects the exact-fp32 CUDA-core kernel (self-check path).
 */
XP_API int64_t xp_match_workspace_bytes(int64_t P, int64_t n1_stride, int64_t n2_stride, int64_t C);
XP_API int xp_mnn_match(const float* d1, const float* d2, const int32_t* n1, const int32_t* n2, int64_t P,
                 int64_t n1_stride, int64_t n2_stride, int64_t C, int32_t* nn12, int32_t* nn21, int32_t* match_idx,
                 float* match_dist, int32_t* match_count, int32_t use_tensor_cores, void* workspace,
                 int64_t workspace_bytes, xp_stream_t stream);

/* -- f3 ("next" row): homography from the mutual matches -----------------------------------------
 * Replaces the per-pair CPU call cv2.findHomography(optical_pts, thermal_pts, USAC_MAGSAC, ransacReprojThreshold,
 * confidence 0.9999, maxIters 10000) of the evaluation (xpoint/utils/evaluation.py:359-378; points are (x, y) =
 * (kp[1], kp[0]), H maps image-1 points onto image-2 points) for the whole batch, without leaving the GPU.
 * OpenCV's USAC is randomised and version-dependent, so the contract is H and the inlier set, not its internals: this is a
 * deterministic LO-RANSAC (hash-sampled 4-point DLT hypotheses, forward transfer error < reproj_threshold, most inliers /
 * lowest index wins, lo_rounds least-squares refits over the inliers), fp64, restated line by line in the oracle.
 *   kp1, kp2 (B, k, 2) int32 (y, x); n1 (B) int32 valid rows of kp1 (NULL = k); match_idx (B, k) int32: row of kp2 matched
 *   to kp1 row i, or -1.  H (B, 3, 3) fp64 row-major with H[2][2] = 1; inlier_mask (B, k) u8 per kp1 row (nullable);
 *   n_inliers (B) int32, -1 when fewer than 4 matches / no valid hypothesis (the reference's H_est = None; H is 0 then).
 *   At most 11 264 matches per pair take part (the first ones in kp1 order); XPoint's 4 096 / 16 384-keypoint configurations
 *   stay below that.
 */
XP_API int xp_estimate_homography(const int32_t* kp1, const int32_t* kp2, const int32_t* n1, const int32_t* match_idx,
                                  int64_t B, int64_t k, int64_t height, int64_t width, int32_t iters, float reproj_threshold,
                                  int32_t lo_rounds, uint32_t seed, double* H, uint8_t* inlier_mask, int32_t* n_inliers,
                                  xp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* XPOINT_B200_H_ */
