/*
 * arflow_b200 — C-ABI boundary of the B200-native unsupervised-flow hot path.
 *
 * Every entry point takes raw DEVICE pointers (fp32, contiguous NCHW unless stated), explicit
 * sizes, and the CUDA stream to launch on (a cudaStream_t passed as void*).  Functions never
 * allocate, never synchronise and never throw; outputs and workspaces are allocated by the
 * caller.  Return value: 0 = ok, <0 = argument error (ARF_E*), >0 = cudaError_t of the launch.
 * The current CUDA device/context of the calling thread is used (the Python shim calls inside
 * torch's device guard, like the reference does with torch.cuda.device_of, correlation.py:21).
 *
 * Reference interfaces replaced (paths relative to deu439/ARFlow):
 *   arf_corr_*      correlation_cuda.forward/backward  models/correlation_package/correlation_cuda.cc:10-16,89-96,169-172
 *                   Correlation.forward (native)       models/correlation_native.py:13-23
 *                   compute_cost_volume                models/uflow_model.py:53-92
 *   arf_warp_*      flow_warp -> F.grid_sample         utils/warp_utils.py:83-90
 *                   resample  -> F.grid_sample         utils/uflow_utils.py:53-77
 *   arf_resampler_* resampler_with_unstacked_warp      utils/uflow_resampler.py:155-241
 *   arf_range_map   compute_range_map                  utils/uflow_utils.py:80-160, utils/warp_utils.py:158-239
 *   arf_corr_map    get_corresponding_map              utils/warp_utils.py:26-80
 *   arf_mask_*      mask_invalid / border_mask         utils/uflow_utils.py:35-50, utils/warp_utils.py:119-134
 *   arf_occ_bidir   get_occu_mask_bidirection          utils/warp_utils.py:93-100
 *   arf_census_*    census_loss(_no_penalty), TernaryLoss   utils/uflow_utils.py:241-306, losses/loss_blocks.py:12-62
 *   arf_ssim_*      ssim_loss, SSIM                    utils/uflow_utils.py:309-334, losses/loss_blocks.py:65-84
 *   arf_smooth_*    UFlowLoss smoothness block, smooth_loss_no_penalty, smooth_grad_1st/2nd
 *                                                      losses/uflow_loss.py:58-102, losses/uflow_elbo_loss.py:81-96, losses/loss_blocks.py:87-124
 *   arf_resize_*    upsample / downsample (bilinear, align_corners=False)   utils/uflow_utils.py:163-204
 *   arf_stencil_mv_*  matrix_vector_product(_T)_general    utils/triag_solve.py:29-43,59-73
 *   arf_trisolve    triag_solve_cuda.forward/backward_substitution   utils/triag_solve/triag_solve.cpp:12-36
 *   arf_inv_diag    triag_solve_cuda.inverse_diagonal  utils/triag_solve/triag_solve.cpp:38-45
 *   arf_featnorm_*  normalize_features                 models/uflow_model.py:8-50
 */
#ifndef ARFLOW_B200_H
#define ARFLOW_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARF_OK            0
#define ARF_EINVAL       -1   /* bad shape / parameter / null pointer */
#define ARF_EUNSUPPORTED -2   /* valid in the reference but not implemented here */
#define ARF_EWORKSPACE   -3   /* workspace too small */
#define ARF_ETIMEOUT     -4   /* a peer did not reach an all-reduce barrier within ~2 s */

/* padding modes of the warp (grid_sample padding_mode) */
#define ARF_PAD_ZEROS      0
#define ARF_PAD_BORDER     1
#define ARF_PAD_REFLECTION 2
/* interpolation modes */
#define ARF_INTERP_BILINEAR 0
#define ARF_INTERP_NEAREST  1
/* what the 2-channel field holds */
#define ARF_FIELD_FLOW   0    /* displacement: sample at (j + u, i + v)            (flow_warp) */
#define ARF_FIELD_COORDS 1    /* absolute pixel coordinates (x, y)                 (resample)  */

int         arf_version(void);
const char* arf_error_string(int code);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
long long   arf_launch_count(void);
/* Test hook. key 0: value != 0 forces the non-TMA (cp.async) staging path of the tiled kernels; keys 1-2 pick
 * correlation variants, key 3 warp variants (1 direct, 2 window), key 4 value 1 forces the wavefront solve. */
int         arf_debug_set(int key, int value);

/* ---------------------------------------------------------------- correlation ---------- */
/* Output dims of the cost volume, same arithmetic as correlation_cuda.cc:25-34. */
int arf_corr_out_dims(int H, int W, int pad, int ks, int md, int s1, int s2,
                      int* D2, int* oH, int* oW);

/* out[b,(tj+dr)*D+(ti+dr),oy,ox] = 1/(ks*ks*C) * sum_{j,i,c} f1p[b,c,y1+j,x1+i] * f2p[b,c,y1+tj*s2+j,x1+ti*s2+i]
 * with y1 = oy*s1+md, x1 = ox*s1+md in the zero-padded frame (pad on every side), dr = md/s2, D = 2dr+1
 * (correlation_cuda_kernel.cu:41-114).  f1,f2: (B,C,H,W); out: (B,D*D,oH,oW). */
int arf_corr_fwd(const float* f1, const float* f2, float* out,
                 int B, int C, int H, int W,
                 int pad, int ks, int md, int s1, int s2, void* stream);

/* gout: (B,D*D,oH,oW); g1,g2: (B,C,H,W), fully overwritten (correlation_cuda_kernel.cu:116-300).
 * g1 or g2 may be NULL to skip that gradient. */
int arf_corr_bwd(const float* f1, const float* f2, const float* gout, float* g1, float* g2,
                 int B, int C, int H, int W,
                 int pad, int ks, int md, int s1, int s2, void* stream);

/* ---------------------------------------------------------------- backward warp -------- */
/* y[b,c,i,j] = interp(x[b,c], X, Y) with, in fp32 and in this order (the reference's normalise ->
 * grid_sample un-normalise round trip, warp_utils.py:16-23 / uflow_utils.py:72-76):
 *   px = field==FLOW ? float(j) + f[b,0,i,j] : f[b,0,i,j]      (same for py with i, channel 1)
 *   gx = 2*px/nW1 - 1,  gy = 2*py/nH1 - 1          nW1,nH1 = normalisation divisors (W-1, H-1, or max(.,1))
 *   X  = align ? (gx+1)/2*(Ws-1) : ((gx+1)*Ws-1)/2 (ATen grid_sampler_unnormalize)
 * x: (B,C,Hs,Ws); field: (B,2,Ho,Wo); y: (B,C,Ho,Wo). */
int arf_warp_fwd(const float* x, const float* field, float* y,
                 int B, int C, int Hs, int Ws, int Ho, int Wo,
                 float nW1, float nH1, int field_kind, int interp, int pad_mode, int align_corners,
                 void* stream);

/* gy: (B,C,Ho,Wo).  gx (B,C,Hs,Ws) may be NULL (source detached, e.g. uflow_loss.py:31,34); when given it
 * is zero-filled here and accumulated with atomics.  gfield (B,2,Ho,Wo) may be NULL. */
int arf_warp_bwd(const float* x, const float* field, const float* gy, float* gx, float* gfield,
                 int B, int C, int Hs, int Ws, int Ho, int Wo,
                 float nW1, float nH1, int field_kind, int interp, int pad_mode, int align_corners,
                 void* stream);

/* ---------------------------------------------------------------- masks / range map ---- */
/* strict=0: mask_invalid (uflow_utils.py:35-50), 1[0<=x<=W-1 && 0<=y<=H-1];
 * strict=1: border_mask (warp_utils.py:119-134), 1[0<x<W-1 && 0<y<H-1].  field (B,2,H,W) -> mask (B,1,H,W). */
int arf_inside_mask(const float* field, float* mask, int B, int H, int W, int field_kind, int strict, void* stream);

/* compute_range_map (uflow_utils.py:80-160, warp_utils.py:158-239; field_kind FLOW) and
 * get_corresponding_map (warp_utils.py:26-80; field_kind COORDS): bilinear forward splat count.
 * count (B,1,H,W) is zero-filled here. */
int arf_range_map(const float* field, float* count, int B, int H, int W, int field_kind, void* stream);
/* gradient of the splat w.r.t. the field (needed where the reference does not detach the range map:
 * the occlusion penalty of uflow_elbo_loss.py:552-559). */
int arf_range_map_bwd(const float* field, const float* gcount, float* gfield, int B, int H, int W,
                      int field_kind, void* stream);

/* mode 0: clamp(c,0,1)   1: clamp(c,0,1) < th   2: 1 - clamp(c,0,1)     (uflow_loss.py:41, warp_utils.py:111-116) */
int arf_count_to_mask(const float* count, float* out, long long n, int mode, float th, void* stream);

/* get_occu_mask_bidirection tail (warp_utils.py:93-100): out = |f12+f21w|^2 > scale*(|f12|^2+|f21w|^2)+bias */
int arf_occ_bidir(const float* flow12, const float* flow21_warped, float* out, int B, int H, int W,
                  float scale, float bias, void* stream);

/* ---------------------------------------------------------------- bilinear resize ------ */
/* F.interpolate(mode='bilinear'): align_corners=0 as used by upsample/downsample (uflow_utils.py:163-204), rh, rw =
 * source step per destination pixel (= 1/scale_factor); align_corners=1 as used by the PWC-Lite flow up-sampling
 * (models/pwclite.py:178-179, 203), rh = (Hi-1)/(Ho-1), rw = (Wi-1)/(Wo-1).  out = mul * interp (mul scales flow
 * values).  in: (planes,Hi,Wi) -> out: (planes,Ho,Wo).  The backward is a deterministic gather. */
int arf_resize_bilinear_fwd(const float* in, float* out, long long planes, int Hi, int Wi, int Ho, int Wo,
                            float rh, float rw, float mul, int align_corners, void* stream);
int arf_resize_bilinear_bwd(const float* gout, float* gin, long long planes, int Hi, int Wi, int Ho, int Wo,
                            float rh, float rw, float mul, int align_corners, void* stream);

/* ---------------------------------------------------------------- census / ternary ----- */
/* Number of per-CTA partial sums (2 floats each) the fused reduction needs for a (B,3,H,W) image pair. */
int arf_census_num_partials(int B, int H, int W);

/* hamming[b,0,y,x] = scale * sum_k sq_k/(0.1+sq_k), sq_k = (t_a,k - t_b,k)^2, t = d/sqrt(0.81+d^2),
 * d = gray255(neighbour k) - gray255(centre), zero outside the image (census_transform + soft_hamming,
 * uflow_utils.py:241-279; TernaryLoss, loss_blocks.py:12-62; scale = 1 for the sum, 1/patch^2 for the mean).
 * If sums != NULL also: sums[0] = sum((|h|+eps)^q * pm), sums[1] = sum(pm), sums[2] = sums[0]/(sums[1]+1e-6)
 * with pm = mask with a patch/2 border zeroed (mask NULL = ones) — census_loss, uflow_utils.py:282-293.
 * im_a, im_b: (B,3,H,W); mask, hamming: (B,1,H,W); partials: 2*arf_census_num_partials floats. patch in {3,5,7}. */
int arf_census_fwd(const float* im_a, const float* im_b, const float* mask, float* hamming,
                   float* partials, float* sums, int B, int H, int W, int patch, float scale,
                   float eps, float q, void* stream);

/* Gradients w.r.t. the RGB images (either may be NULL).  Upstream gradient: ghamming (B,1,H,W) if given,
 * else the fused census_loss tail is differentiated from (hamming, mask, sums, gloss[0]). */
int arf_census_bwd(const float* im_a, const float* im_b, const float* ghamming, const float* hamming,
                   const float* mask, const float* sums, const float* gloss, float* g_a, float* g_b,
                   int B, int H, int W, int patch, float scale, float eps, float q, void* stream);

/* The same for `groups` equally sized, consecutive slices of the batch in ONE launch each way (B % groups == 0): every
 * slice is its own census_loss - own mask sum and normaliser (uflow_utils.py:293) - with sums[3g + {0,1,2}] and
 * gloss[g] for slice g.  UFlowLoss (losses/uflow_loss.py:28-54) evaluates its two directions this way, stacked on the
 * batch.  arf_census_fwd / arf_census_bwd are the groups = 1 case. */
int arf_census_fwd_groups(const float* im_a, const float* im_b, const float* mask, float* hamming,
                          float* partials, float* sums, int B, int H, int W, int groups, int patch, float scale,
                          float eps, float q, void* stream);
int arf_census_bwd_groups(const float* im_a, const float* im_b, const float* ghamming, const float* hamming,
                          const float* mask, const float* sums, const float* gloss, float* g_a, float* g_b,
                          int B, int H, int W, int groups, int patch, float scale, float eps, float q, void* stream);

/* ---------------------------------------------------------------- smoothness ----------- */
int arf_smooth_num_partials(int B, int H, int W);

/* out[0] = final_scale * (mean_x(w_x*pen(d_x)) + mean_y(w_y*pen(d_y))), see csrc/smooth.cu for the
 * parameter table (uflow_loss.py:58-102, loss_blocks.py:93-124).  img: (B,Ci,H,W), flow: (B,2,H,W).
 * penalty 0: sqrt(d^2+eps2), 1: |d|.  partials: 2*arf_smooth_num_partials floats. */
int arf_smooth_fwd(const float* img, const float* flow, float* out, float* partials, int B, int Ci, int H, int W,
                   int order, int wstride, int woff, int penalty, float edge, float eps2, float final_scale,
                   void* stream);
int arf_smooth_bwd(const float* img, const float* flow, const float* gloss, float* gflow, int B, int Ci, int H, int W,
                   int order, int wstride, int woff, int penalty, float edge, float eps2, float final_scale,
                   void* stream);

/* ---------------------------------------------------------------- feature normalisation */
/* normalize_features([f1, f2], normalize=True, center=True, moments_across_channels=True, moments_across_images=True)
 * (models/uflow_model.py:8-50; the only setting the models use, :163-170): per sample, mean and unbiased variance
 * over all n = C*H*W elements of each map, averaged over the two maps; y_k = (f_k - mu) / sqrt(var + 1e-16).
 * f*, y*, g*, d*: (B, n) dense; stats: B*4 floats (mu, std, mu_1, mu_2) written by fwd and read by bwd;
 * coef: B*2 floats scratch; ws: arf_featnorm_workspace(B, n) bytes.  d1 / d2 may be NULL. */
long long arf_featnorm_workspace(long long B, long long n);
int arf_featnorm_fwd(const float* f1, const float* f2, float* y1, float* y2, float* stats, void* ws, long long B,
                     long long n, void* stream);
int arf_featnorm_bwd(const float* f1, const float* f2, const float* g1, const float* g2, const float* stats, float* d1,
                     float* d2, float* coef, void* ws, long long B, long long n, void* stream);

/* ---------------------------------------------------------------- conv epilogue -------- */
/* Bias + leaky ReLU around the (cuDNN) convolutions of the PWC networks: nn.Conv2d(bias=True) followed by
 * nn.LeakyReLU / func.leaky_relu (models/uflow_model.py:134-135, 427-436; uflow_prob_model.py:445-456).
 * fwd, in place on the convolution output y (B,C,HW):  y = leaky(y + bias[c]); bias may be NULL.
 * bwd, one pass:  g = gy * (y > 0 ? 1 : slope)  and  dbias[c] = sum_{b,hw} g  (dbias NULL: no reduction).
 * partials: arf_bias_leaky_num_partials(B,C,HW) floats (two-stage, deterministic bias gradient). */
long long arf_bias_leaky_num_partials(long long B, int C, long long HW);
int arf_bias_leaky_fwd(float* y, const float* bias, long long B, int C, long long HW, float slope, void* stream);
int arf_bias_leaky_bwd(const float* gy, const float* y, float* g, float* partials, float* dbias, long long B, int C,
                       long long HW, float slope, void* stream);

/* Channels-last (NHWC) variants: y, gy, g are (rows = N*H*W) x C row-major; partials: arf_bias_leaky_nhwc_num_partials. */
long long arf_bias_leaky_nhwc_num_partials(long long rows, int C);
int arf_bias_leaky_nhwc_fwd(float* y, const float* bias, long long rows, int C, float slope, void* stream);
int arf_bias_leaky_nhwc_bwd(const float* gy, const float* y, float* g, float* partials, float* dbias, long long rows,
                            int C, float slope, void* stream);
/* same, with gy and y column slices of wider row-major matrices (row strides gy_ld, y_ld >= C): the dense block's
 * backward reads a layer's output gradient straight out of the gradient of the concatenation it went into, and the
 * layer's activated output out of that concatenation itself */
int arf_bias_leaky_nhwc_bwd_ld(const float* gy, long long gy_ld, const float* y, long long y_ld, float* g, float* partials,
                               float* dbias, long long rows, int C, float slope, void* stream);
/* forward with a separate destination: dst[r*dst_ld + c] = leaky(src[r*C + c] + bias[c]) — the activated output lands
 * directly in its column slice of the next dense-block input (dst_ld = that input's width) */
int arf_bias_leaky_nhwc_fwd_ld(const float* src, float* dst, long long dst_ld, const float* bias, long long rows, int C,
                               float slope, void* stream);

/* The flow-output convolutions, nn.Conv2d(Cin, 2, 3, padding=1) at the end of every decoder level and of the
 * refinement network (models/uflow_model.py:139-143, 232-249), in fp32 on a channels-last input.
 * x: (N,H,W,Cin) packed channels-last, Cin % 32 == 0; w: (2,3,3,Cin) = a channels-last (2,Cin,3,3) weight;
 * y / gy: (N,2,H,W) NCHW.  fwd: y = conv(x, w) + bias (bias may be NULL).
 * bwd: gx (N,H,W,Cin) input gradient (NULL: skipped, then w may be NULL too); out: 2*9*Cin floats dW in
 * (co, kh, kw, ci) order, then 2 floats dbias; partials: arf_conv3x3_small_bwd_workspace(...) floats.
 * Other Cout / Cin: ARF_EUNSUPPORTED. */
int arf_conv3x3_small_fwd(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int Cin,
                          int Cout, void* stream);
long long arf_conv3x3_small_bwd_workspace(int N, int H, int W, int Cin, int Cout);
int arf_conv3x3_small_bwd(const float* x, const float* gy, const float* w, float* gx, float* out, float* partials, int N,
                          int H, int W, int Cin, int Cout, void* stream);

/* Weight gradient of the first pyramid convolution, nn.Conv2d(3, 32, 3, stride=2, padding=1) (models/uflow_model.py:
 * 427-436), on the 8-channel channels-last image (3 real channels, 5 zeros).  x: (N,Hi,Wi,8); g: (N,Ho,Wo,32)
 * channels-last gradient of the convolution output, Ho = (Hi-1)/2+1; out: 27*32 floats in [kh][kw][ci][co] order;
 * partials: arf_conv3x3s2_first_wgrad_workspace(N,Hi,Wi) floats.  Other channel counts: ARF_EUNSUPPORTED. */
long long arf_conv3x3s2_first_wgrad_workspace(int N, int Hi, int Wi);
int arf_conv3x3s2_first_wgrad(const float* x, const float* g, float* out, float* partials, int N, int Hi, int Wi,
                              int Cin_real, int Cout, void* stream);

/* ---------------------------------------------------------------- NHWC concat ---------- */
/* The decoder's torch.cat([...], dim=1) (models/uflow_model.py:189-205) into a packed NHWC tensor of Cd channels
 * (Cd >= sum of the parts, the tail is padding): pack writes one part at channel offset c_off; src is NHWC
 * (N,HW,Cs) when src_nhwc, else NCHW (N,Cs,HW); src == NULL writes zeros (padding channels).  unpack is the
 * inverse (backward of the concat): part <- packed[..., c_off : c_off+Cs] in the part's own layout. */
int arf_nhwc_pack(float* dst, const float* src, long long N, long long HW, int Cs, int Cd, int c_off, int src_nhwc,
                  void* stream);
/* Convolution weight (Co,Ci,KH,KW; element strides s_*) <-> its channels-last copy (Co_pad,Ci_pad,KH,KW) with
 * pad_cnt[k] zero input channels inserted at original position pad_at[k] (k < n_pads <= 4, increasing) and zero output
 * channels appended.  to_padded = 1: dst = padded copy; 0: dst = gradient in the parameter's layout, src = padded. */
int arf_pad_weight(float* dst, const float* src, int Co, int Ci, int KH, int KW, int Co_pad, int Ci_pad, long long s_co,
                   long long s_ci, long long s_kh, long long s_kw, int n_pads, const int* pad_at, const int* pad_cnt,
                   int to_padded, void* stream);
int arf_nhwc_unpack(float* part, const float* packed, long long N, long long HW, int Cs, int Cd, int c_off,
                    int part_nhwc, void* stream);
/* pack / unpack of an NCHW part with a leaky ReLU fused into the copy (the cost volume's activation,
 * models/uflow_model.py:185-186): dst slice = leaky(src);  part = packed_grad * (packed_fwd > 0 ? 1 : slope). */
int arf_nhwc_pack_act(float* dst, const float* src, long long N, long long HW, int Cs, int Cd, int c_off, float slope,
                      void* stream);
int arf_nhwc_unpack_act(float* part, const float* packed_grad, const float* packed_fwd, long long N, long long HW, int Cs,
                        int Cd, int c_off, float slope, void* stream);
/* NCHW (N,C,HW) <-> channels-last (N,HW,C) copy through a tiled transpose, the NCHW side's batch index rotated by
 * batch_shift (the stacked flow directions read each other's features: `feature_pyramid2` of uflow_model.py:255-257):
 * to_nchw = 1: nchw[(n+shift) % N] = nhwc[n];  to_nchw = 0: nhwc[n] = nchw[(n+shift) % N]. */
int arf_nhwc_transpose(float* dst, const float* src, long long N, long long HW, int C, int to_nchw, int batch_shift,
                       void* stream);
/* part (NHWC) += packed[..., c_off : c_off+Cs]: the gradient of a tensor that feeds both a convolution and the next
 * concatenation is accumulated in place instead of unpack + add */
int arf_nhwc_unpack_add(float* part, const float* packed, long long N, long long HW, int Cs, int Cd, int c_off,
                        void* stream);

/* ---------------------------------------------------------------- stencil-triangular ---- */
/* matrix_vector_product_general / _T_general (utils/triag_solve.py:29-43, 59-73).
 * A: (N, 2*(k+1)^2, H, W), tap t = i*(k+1)+j occupies channels 2t, 2t+1 (u, v); X, Y: (N,2,H,W).
 *   transposed=0:  Y[p] = sum_t A[t][p-(i,j)] * X[p-(i,j)]      (y = L x, taps leaving the image dropped)
 *   transposed=1:  Y[p] = sum_t A[t][p] * X[p+(i,j)]            (y = L^T x) */
int arf_stencil_mv_fwd(const float* A, const float* X, float* Y, int N, int H, int W, int k, int transposed,
                       void* stream);
/* dA (same shape as A) and dX (same shape as X) of the product above; either may be NULL. */
int arf_stencil_mv_bwd(const float* A, const float* X, const float* gY, float* dA, float* dX, int N, int H, int W,
                       int k, int transposed, void* stream);

/* forward_substitution (upper=0) / backward_substitution (upper=1) of triag_solve_cuda
 * (utils/triag_solve/triag_solve.cpp:12-36, triag_solve_cuda.cu:7-69; Python twins triag_solve.py:76-115).
 * A: (S,M,N) diagonal, B: (S,M,N-1) left/right, C: (S,M-1,N) above/below, D: (S,M-1,N-1) diagonal neighbour
 * (may be NULL), X -> Y: (S,M,N); S = batch*channels systems.  Y may not alias X.  N <= 1024 runs the row-scan
 * kernel, wider systems the anti-diagonal wavefront (M <= 1024, else ARF_EUNSUPPORTED). */
int arf_trisolve(const float* A, const float* B, const float* C, const float* D, const float* X, float* Y,
                 long long systems, int M, int N, int upper, void* stream);

/* inverse_diagonal (triag_solve.cpp:38-45, triag_solve_cuda.cu:72-139): H[s,k,l] = || L^-1 e_(k,l) ||^2,
 * L built from A, B, C only. */
int arf_inv_diag(const float* A, const float* B, const float* C, float* H, long long systems, int M, int N,
                 void* stream);

/* ---------------------------------------------------------------- SSIM ------------------ */
/* x, y: (planes,H,W).  valid=0: zero-padded P x P box filters with divisor P*P, output (planes,H,W)
 * (ssim_loss, uflow_utils.py:309-334); valid=1: unpadded filters, output (planes,H-P+1,W-P+1)
 * (SSIM, loss_blocks.py:65-84).  mode 0: out1 = clamp(1-S1,0,1), out2 = clamp(1-S2,0,1);
 * mode 1: out1 = clamp((1 - S1*S2)/2, 0, 1), out2 unused.  patch in {3,5,7}. */
int arf_ssim_fwd(const float* x, const float* y, float* out1, float* out2, long long planes, int H, int W,
                 int patch, int valid, int mode, void* stream);
/* g1, g2: upstream gradients of out1, out2; coef: workspace of 5*planes*Ho*Wo floats; gx, gy may be NULL. */
int arf_ssim_bwd(const float* x, const float* y, const float* g1, const float* g2, float* coef, float* gx,
                 float* gy, long long planes, int H, int W, int patch, int valid, int mode, void* stream);

/* ---------------------------------------------------------------- NHWC resampler -------- */
/* resampler_with_unstacked_warp(data, warp_x, warp_y, safe=True) (utils/uflow_resampler.py:155-241).
 * data: (B,H,W,C); warp_x/warp_y: B*P coordinates read with element stride wstride (2 for the interleaved
 * (...,2) tensor of `resampler`, :137-152); out: (B,P,C).  Taps floor/ceil, out-of-range taps contribute 0. */
int arf_resampler_fwd(const float* data, const float* warp_x, const float* warp_y, long long wstride, float* out,
                      int B, int H, int W, int C, long long P, void* stream);
/* gdata (zero-filled here, atomics), gwx/gwy written with element stride gwstride; each may be NULL. */
int arf_resampler_bwd(const float* data, const float* warp_x, const float* warp_y, long long wstride,
                      const float* gout, float* gdata, float* gwx, float* gwy, long long gwstride,
                      int B, int H, int W, int C, long long P, void* stream);

/* Input stage of the stacked-direction networks: src (B, 2C, HW) image pairs -> dst (2B, HW, 8) channels-last,
 * batch order [first images; second images], value * scale + shift on the C real channels, zeros behind them
 * (torch.cat of the two slices, x * 2 - 1 and the 3 -> 8 channel pack of models/uflow_model.py:404, 199-205).
 * Cd must be 8 and C <= 8, else ARF_EUNSUPPORTED. */
int arf_image_pair_pack(float* dst, const float* src, long long B, long long HW, int C, int Cd, float scale, float shift,
                        void* stream);

/* ---------------------------------------------------------------- gradient all-reduce ---- */
/* Data-parallel gradient exchange over NVLink peer memory (one process per GPU); replaces, for the benchmark driver,
 * nn.DataParallel's gradient gather of the reference (trainer/base_trainer.py:75,131-147).
 * arf_comm_alloc: cudaMalloc + zero fill of an IPC-exportable buffer (gradient storage, or a flag area of
 * arf_comm_flag_bytes() bytes).  arf_comm_ipc_get / _open / _close: 64-byte CUDA IPC handle of such a buffer, and its
 * mapping in a peer process (peer access is enabled on first use). */
int arf_comm_flag_bytes(void);
int arf_comm_alloc(void** ptr, size_t bytes);
int arf_comm_free(void* ptr);
int arf_comm_ipc_get(void* ptr, void* handle64);
int arf_comm_ipc_open(const void* handle64, void** ptr);
int arf_comm_ipc_close(void* ptr);
/* In-place all-reduce of data[rank][offset .. offset+count) (floats; offset, count multiples of 4) across nranks <= 8
 * processes: data[p] / flags[p] are rank p's gradient buffer and flag area as mapped in THIS process (host arrays of
 * device pointers).  Result = scale * sum over ranks, bit-identical on every rank.  One kernel launch of `ctas` (<= 64)
 * CTAs, no host synchronisation: capturable in a CUDA graph; every rank must issue the same sequence of calls with the
 * same ctas. */
int arf_allreduce_f32(float* const* data, unsigned* const* flags, int rank, int nranks, size_t offset, size_t count,
                      float scale, int ctas, void* stream);
/* Synchronises the stream; ARF_ETIMEOUT if a barrier of an earlier arf_allreduce_f32 on this flag area gave up. */
int arf_comm_error(const unsigned* flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARFLOW_B200_H */
