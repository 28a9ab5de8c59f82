/* b200eg3d.h -- C ABI of libb200eg3d.so: the sm_100a kernels behind the EG3D tri-plane generator hot path
 * (TriPlaneGenerator.synthesis of cvlab-kaist/3DGAN-Inversion).
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; all pointers are DEVICE pointers of the current CUDA device unless noted
 *   - nothing is allocated, nothing synchronises: the caller owns inputs, outputs and workspaces.  The only process-wide
 *     state are the three tuning switches below (b200_set_*) and per-DEVICE one-time launch attributes / SM counts (keyed by
 *     cudaGetDevice(), so several GPUs can be driven from one process); entry points are re-entrant across streams and devices
 *   - the CALLER makes the pointers' device current (cudaSetDevice) before the call -- the Python wrapper does so per call
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and the call returns immediately
 *   - returns 0 on success; non-zero on error, with a message available from b200_last_error() (thread-local)
 *   - activations are NHWC fp32: x[n][h][w][c]; modulated weights are wmod[n][tap][cout][cin], tap = kh*ksize + kw
 *   - "bf16" pointers are raw __nv_bfloat16 arrays in the same layouts (split-float operands hi + lo ~= fp32 value)
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference repository).
 */
#ifndef B200EG3D_H
#define B200EG3D_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------------ */
int b200_version(void);                 /* replaces the plugin identity of torch_utils/custom_ops.py:61 get_plugin() */
const char* b200_last_error(void);      /* replaces TORCH_CHECK -> RuntimeError text, torch_utils/ops/bias_act.cpp:39-55 */
int b200_set_pdl(int on);               /* programmatic dependent launch of the conv / epilogue / FIR kernels on (default; env
                                           B200EG3D_PDL=0 disables) or off; returns the previous setting.  Profiling aid. */

int b200_set_conv_pair(int on);         /* CTA-pair (cta_group::2, M = 256) tiles of the tensor-core forward / dgrad convolutions on (default;
                                           env B200EG3D_CONV_PAIR=0 disables) or off = the single-CTA kernel, kept as the cross-check.
                                           Returns the previous setting. */
int b200_set_mlp_passes(int passes);    /* operand passes of the fused decoder MLP forward: 3 = split operands, fp32-equivalent (default;
                                           env B200EG3D_MLP_PASSES=1 selects 1), 1 = single pass ("fast mode", non-parity).
                                           Returns the previous setting. */

/* ---- modulated convolution (training/networks_stylegan2.py:34-91 modulated_conv2d, fused path) ---------------- */

/* w' = W * s; d = rsqrt(sum w'^2 + 1e-8); wmod = w' * d (demod != 0), written in GEMM layout.  networks_stylegan2.py:58-67.
 * W [cout][cin][taps], styles [n][cin], wmod [n][taps][cout][cin], dcoef [n][cout] (may be NULL when demod == 0). */
int b200_modconv_weight_prep(const float* W, const float* styles, float* wmod /* fp32, may be NULL */,
                             void* w_hi_bf16 /* may be NULL */, void* w_lo_bf16 /* may be NULL */, float* dcoef,
                             int n, int cout, int cin, int taps, int demod, void* stream);
/* Backward of the above: dwmod -> dW [cout][cin][taps] (overwritten), dstyles [n][cin] (overwritten; may be NULL). */
int b200_modconv_weight_prep_bwd(const float* W, const float* styles, const float* dcoef, const float* dwmod,
                                 float* dW, float* dstyles, int n, int cout, int cin, int taps, int demod, void* stream);

/* Grouped ("bank") style + weight preparation: every modulated-conv layer of a synthesis network in one launch per stage.
 * Replaces the per-layer `styles = self.affine(w)` (networks_stylegan2.py:312,354) and the weight modulation of :58-67 for
 * up to 32 layers at once.  `layers` is a HOST array (copied into the kernel parameters); all pointers inside are device
 * pointers.  post_scale multiplies the affine output (ToRGB's weight_gain, networks_stylegan2.py:354); demod selects :65-67. */
#define B200_BANK_MAX_LAYERS 32
typedef struct {
    const float* affine_w;      /* [cin][w_dim] */
    const float* affine_b;      /* [cin] */
    const float* weight;        /* [cout][cin][taps] */
    float* styles;              /* out [n][cin] (post-scaled) */
    float* dcoef;               /* out [n][cout], demod only */
    float* wmod;                /* out fp32 [n][taps][cout][cin] or NULL */
    void* w_hi; void* w_lo;     /* out split bf16, same layout, or NULL */
    const float* dwmod;         /* backward in: gradient w.r.t. wmod, or NULL (layer skipped) */
    float* d_weight;            /* backward out [cout][cin][taps] or NULL */
    float* d_styles;            /* backward scratch [n][cin], zeroed by the caller, or NULL */
    float* d_affine_w; float* d_affine_b;   /* backward out or NULL */
    int widx, cin, cout, taps, demod;
    float post_scale;
} B200BankLayer;
int b200_bank_styles_fwd(const B200BankLayer* layers, int n_layers, const float* ws /* [n][num_ws][w_dim] */, int n, int num_ws,
                         int w_dim, void* stream);
int b200_bank_weights_fwd(const B200BankLayer* layers, int n_layers, int n, void* stream);
int b200_bank_weights_bwd(const B200BankLayer* layers, int n_layers, int n, void* stream);
int b200_bank_styles_bwd(const B200BankLayer* layers, int n_layers, const float* ws, float* d_ws /* accumulated, may be NULL */,
                         int n, int num_ws, int w_dim, void* stream);

/* Exact-fp32 gather-GEMM convolutions (any channel count).  Replace F.conv2d / F.conv_transpose2d reached through
 * torch_utils/ops/conv2d_gradfix.py:37-45 from torch_utils/ops/conv2d_resample.py:113-136, and their autograd.
 * ksize in {1,3}; up == 1: 'same' correlation (padding ksize/2); up == 2 (ksize 3): stride-2 transposed convolution whose
 * output y / incoming dy live on the (2h+1) x (2w+1) grid (conv2d_resample.py:121-127, before the FIR of :128).
 * h, w are the INPUT spatial size. */
int b200_conv_fwd(const float* x, const float* wmod, float* y, int n, int h, int w, int cin, int cout,
                  int ksize, int up, void* stream);
int b200_conv_dgrad(const float* dy, const float* wmod, float* dx, int n, int h, int w, int cin, int cout,
                    int ksize, int up, void* stream);
int b200_conv_wgrad(const float* x, const float* dy, float* dwmod, int n, int h, int w, int cin, int cout,
                    int ksize, int up, void* stream);
/* The same weight gradient for a 1x1 convolution with <= 4 output channels (the ToRGB layers of the super-resolution blocks) whose
 * input exists only as its split-bf16 pair x = x_hi + x_lo ([n][npix][cin] bf16 each); b200_conv1x1_thin_supported -> 1 if handled. */
int b200_conv1x1_thin_supported(int cin, int cout);
int b200_conv1x1_wgrad_split(const void* x_hi, const void* x_lo, const float* dy, float* dwmod, int n, long npix, int cin, int cout,
                             void* stream);
/* The forward of such a layer from the split pair with the ToRGB bias and clamp (networks_stylegan2.py:353-357) applied on the way
 * out: y [n][npix][cout] = clamp((x_hi + x_lo) . wmod^T + bias, +-clamp).  wmod [n][cout][cin] fp32, bias [cout] or NULL, clamp < 0:
 * none.  b200_conv1x1_fwd_thin_supported -> 1 if handled (cout <= 4, cin a power of two in 8 .. 512). */
int b200_conv1x1_fwd_thin_supported(int cin, int cout);
int b200_conv1x1_fwd_thin(const void* x_hi, const void* x_lo, const float* wmod, const float* bias, float* y, int n, long npix, int cin,
                          int cout, float clamp, void* stream);

/* tcgen05 / TMA tensor-core convolutions on split-bf16 operands (same geometry as above).
 * npass = 3: hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (fp32-parity mode); npass = 1: hi*hi only (lo may be NULL).
 * b200_conv_tc_supported(kind, ...) -> 1 if the shape is handled (kind 0 fwd, 1 dgrad, 2 wgrad), else use the fp32 entry points. */
int b200_conv_tc_supported(int kind, int h, int w, int cin, int cout, int ksize, int up);
int b200_split_bf16(const float* x, void* hi_bf16, void* lo_bf16 /* may be NULL */, long count, void* stream);
/* Layers with few output tiles split K over CTAs and ADD partial sums (red.global): b200_conv_tc_ksplit(kind 0 fwd / 1 dgrad, ...) > 1.
 * Such a launch clears its output first, unless prezeroed != 0: the caller then guarantees an all-zero y / dx (one fill for every
 * split-K output of a network instead of a memset in front of each launch).  Launches that do not split K overwrite either way. */
int b200_conv_tc_ksplit(int kind, int n, int h, int w, int cin, int cout, int ksize, int up);
int b200_conv_fwd_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, float* y,
                     int n, int h, int w, int cin, int cout, int ksize, int up, int npass, int prezeroed, void* stream);
/* up == 1 forward convolution with the SynthesisLayer epilogue (networks_stylegan2.py:318-329) applied while the accumulator
 * leaves tensor memory: z = clamp(lrelu_alpha(conv + noise[pix] * *strength + bias[c]) * act_gain, +-clamp), written as fp32 z (may
 * be NULL: the pair is then the only copy) and as the split-bf16 pair z_hi / z_lo (z_lo may be NULL) that feeds the next convolution.  noise [h*w] (noise_bs 0) or [n][h*w]
 * (noise_bs h*w), may be NULL.  Only for shapes b200_conv_tc_act_fusable reports 1 (no split-K, cout % 32 == 0). */
int b200_conv_tc_act_fusable(int n, int h, int w, int cin, int cout, int ksize);
int b200_conv_fwd_tc_act(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, float* z, void* z_hi, void* z_lo,
                         const float* bias, const float* noise, const float* strength, long noise_bs, int n, int h, int w, int cin,
                         int cout, int ksize, int npass, float alpha, float act_gain, float clamp, void* stream);
int b200_conv_dgrad_tc(const void* dy_hi, const void* dy_lo, const void* w_hi, const void* w_lo, float* dx,
                       int n, int h, int w, int cin, int cout, int ksize, int up, int npass, int prezeroed, void* stream);
/* accumulate == 0: dwmod is overwritten (zeroed inside the call, then reduced into); accumulate != 0: the partial sums are ADDED to
 * dwmod, which the caller zeroed earlier (one fill for every layer of a network, off the critical path). */
int b200_conv_wgrad_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dwmod,
                       int n, int h, int w, int cin, int cout, int ksize, int up, int npass, int accumulate, void* stream);

/* ---- bias / activation (torch_utils/ops/bias_act.cpp:36 bias_act(x,b,xref,yref,dy,grad,dim,act,alpha,gain,clamp)) ---- */

/* Generic op with the plugin's semantics (kernel bias_act.cu:28-151).  act: 1 linear 2 relu 3 lrelu 4 tanh 5 sigmoid 6 elu
 * 7 selu 8 softplus 9 swish.  grad: 0 forward (y = clamp(act(x+b)*gain)), 1 first-order gradient (x is dy).  Any of b, xref,
 * yref, dy may be NULL ("absent", bias_act.cpp:65-68).  Bias index of element i is (i / stepB) % sizeB (bias_act.cu:49). */
int b200_bias_act(const float* x, const float* b, const float* xref, const float* yref, const float* dy, float* y,
                  int grad, long sizeX, long stepB, int sizeB, int act, float alpha, float gain, float clamp, void* stream);

/* SynthesisLayer epilogue on NHWC [n][hw][c]: z = clamp(lrelu(y + noise[pix]*strength + bias[c]) * gain, +-clamp)
 * (training/networks_stylegan2.py:318-329).  noise [hw] (noise_bs 0, shared 'const' buffer) or [n][hw] (noise_bs = hw) or NULL;
 * strength: device scalar; clamp < 0 disables clamping. */
int b200_layer_act_fwd(const float* y, float* z /* may be NULL */, void* z_hi_bf16 /* may be NULL */, void* z_lo_bf16,
                       const float* bias, const float* noise, const float* strength, long noise_bs,
                       int n, int hw, int c, int lrelu, float alpha, float gain, float clamp, void* stream);
/* Backward from the saved OUTPUT (bias_act.cu:73-77,143-145): dy, plus ACCUMULATED dbias[c], dstrength[1], dnoise (any NULL).
 * The output comes as fp32 z or, when z is NULL, as its split-bf16 copy z_hi (+ z_lo, read only where z_hi reaches the clamp). */
int b200_layer_act_bwd(const float* dz, const float* z, const void* z_hi_bf16, const void* z_lo_bf16, float* dy /* may be NULL */,
                       void* dy_hi_bf16 /* may be NULL */, void* dy_lo_bf16, float* dbias, const float* noise, const float* strength,
                       long noise_bs, float* dstrength, float* dnoise, int n, int hw, int c, int lrelu, float alpha, float gain,
                       float clamp, void* stream);
/* The same with the incoming gradient given as two addends dz + dz2: an activation read by the next convolution AND by the block's ToRGB
 * layer (networks_stylegan2.py:449-457) receives two gradients that autograd would first sum in a pass of its own.  Shapes the
 * two-addend form takes: b200_layer_act_bwd_sum2_supported (1 / 0); it needs dy as fp32 or as the full hi + lo pair. */
int b200_layer_act_bwd_sum2(const float* dz, const float* dz2, const float* z, const void* z_hi_bf16, const void* z_lo_bf16,
                            float* dy /* may be NULL */, void* dy_hi_bf16 /* may be NULL */, void* dy_lo_bf16, float* dbias,
                            const float* noise, const float* strength, long noise_bs, float* dstrength, float* dnoise, int n, int hw,
                            int c, int lrelu, float alpha, float gain, float clamp, void* stream);
int b200_layer_act_bwd_sum2_supported(int n, int hw, int c, int has_noise, long noise_bs);

/* ---- upfirdn2d (torch_utils/ops/upfirdn2d.cpp:20 upfirdn2d(x,f,upx,upy,downx,downy,padx0,padx1,pady0,pady1,flip,gain)) ---- */
/* NHWC x [n][h][w][c] -> y [n][oh][ow][c], oh = (h*upy + pady0 + pady1 - fh + downy) / downy (upfirdn2d.cpp:37-38); f [fh][fw];
 * optional `add` (same shape as y) is added to the result (the img.add_(y) of networks_stylegan2.py:457). */
int b200_upfirdn2d(const float* x, const float* f, const float* add, float* y, int n, int h, int w, int c, int fh, int fw,
                   int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                   void* stream);

/* Same filter geometry (square up / down factors) with fused consumers: act != 0 applies the SynthesisLayer epilogue
 * (+ noise*strength + bias, lrelu, act_gain, clamp: conv2d_resample.py:128 followed by networks_stylegan2.py:318-329 in one pass);
 * y (fp32) and y_hi / y_lo (split bf16 for the next tensor-core conv) are optional outputs, at least one is required.
 * separable != 0: the caller guarantees f[a][q] = f[a][0] * f[0][q] / f[0][0] (the [1,3,3,1] x [1,3,3,1] resampling filter of
 * upfirdn2d.setup_filter); 4x4 unit-rate filters then run as a sliding column window (8 instead of 16 MACs per element). */
int b200_upfirdn2d_fused(const float* x, const float* f, const float* add, float* y, void* y_hi_bf16, void* y_lo_bf16,
                         int n, int h, int w, int c, int fh, int fw, int up, int down, int padx0, int padx1, int pady0, int pady1,
                         int flip, float gain, int act, const float* bias, const float* noise, const float* strength,
                         long noise_bs, int lrelu, float alpha, float act_gain, float clamp, int separable, void* stream);

/* ---- fused tri-plane sampling + OSG decoder (renderer.py:39-66 sample_from_planes, triplane.py:124-136 OSGDecoder.forward) ---- */
/* planes [n][hp][wp][96] (plane p = channels 32p..32p+31).  Points: coords [n][P][3], or (coords NULL) rays ray_o/ray_d
 * [n][P/S][3] with depths [n][P] (point p on ray p / S).  ray_w (ray mode; 0 = linear order): width of the ray image -- the
 * tcgen05 kernels then walk 8 x 16-pixel patches front to back (experimental, see tri::map_point; results do not depend on it;
 * a forward and the backward that consumes its f_save must use the same value).  W1 [64][32], b1 [64], W2 [33][64], b2 [33]
 * raw parameters (gains lr_mul/sqrt(fan_in), networks_stylegan2.py:111-112, applied inside).  Out: rgb [n][P][32] (NULL:
 * density-only query), sigma [n][P].
 * f_save (may be NULL): b200_triplane_fsave_bytes(n, P) bytes that receive every 128-point tile's mean features as the finished
 * layer-1 tensor-core operand; handing it to the backward (f_saved) saves the second gather of 12 texel lines per point. */
long b200_triplane_fsave_bytes(int n, long P);
int b200_triplane_mlp_fwd(const float* planes, int n, int hp, int wp, const float* coords, const float* ray_o,
                          const float* ray_d, const float* depths, int S, int ray_w, long P, float box_warp,
                          const float* W1, const float* b1, const float* W2, const float* b2, float lr_mul,
                          float* rgb, float* sigma, void* f_save, void* stream);
/* d_planes is ACCUMULATED (zero it first; may be NULL); d_coords [n][P][3] written (may be NULL); d_ray_o / d_ray_d
 * [n][P/S][3] ACCUMULATED (ray mode; both or neither): the sums over a ray's samples of d point and t * d point, i.e. the
 * gradients of `ray_origins + t * ray_directions` (renderer.py:161,178) w.r.t. origins and directions; dW1..db2 ACCUMULATED
 * (all or none).  dW1 = d_a^T F and dW2 = d_out^T h are contracted over the points inside the kernel (tcgen05.mma into a
 * tensor-memory accumulator).  `workspace`: b200_triplane_bwd_workspace_bytes(n, P) bytes of caller-owned scratch. */
long b200_triplane_bwd_workspace_bytes(int n, long P);
int b200_triplane_mlp_bwd(const float* planes, int n, int hp, int wp, const float* coords, const float* ray_o,
                          const float* ray_d, const float* depths, int S, int ray_w, long P, float box_warp,
                          const float* W1, const float* b1, const float* W2, const float* b2, float lr_mul,
                          const float* d_rgb, const float* d_sigma, const void* f_saved /* from the forward, may be NULL */,
                          float* d_planes, float* d_coords, float* d_ray_o, float* d_ray_d,
                          float* dW1, float* db1, float* dW2, float* db2, void* workspace, long workspace_bytes, void* stream);
/* Implementation switch of the two entry points above: 1 = tcgen05 pipeline (default), 0 = the mma.sync kernels kept as the
 * on-GPU cross-check (env B200EG3D_TRIPLANE_IMPL=0).  Returns the previous setting. */
int b200_set_triplane_impl(int impl);

/* ---- per-ray kernels (renderer.py:143-308 ImportanceRenderer, ray_marcher.py:25-57 MipRayMarcher2) ---------------- */
/* RaySampler.forward (ray_sampler.py:24-73): cam2world [n][16], intrinsics [n][9] -> ray_o, ray_d [n][R*R][3]; the backward
 * ACCUMULATES d cam2world [n][16] (zero it first) from d ray_o / d ray_d (either may be NULL). */
int b200_ray_sampler_fwd(const float* cam2world, const float* intrinsics, int n, int R, float* ray_o, float* ray_d, void* stream);
int b200_ray_sampler_bwd(const float* cam2world, const float* intrinsics, int n, int R, const float* d_ray_o, const float* d_ray_d,
                         float* d_cam2world, void* stream);
/* t[ray][s] = t_base[s] + u[ray][s] * delta          (renderer.py:224-247 sample_stratified, numeric ray_start/ray_end).
 * minmax (may be NULL): the two words of b200_depth_minmax below, updated with the depths this call makes. */
int b200_ray_depths_coarse(const float* t_base, const float* u, float* t, long n_rays, int S, float delta, unsigned* minmax,
                           void* stream);
/* global min / max of depths for the clamp of ray_marcher.py:50; minmax: 2 x uint32 (order-preserving), initialised {~0u, 0} */
int b200_depth_minmax(const float* t, long total, unsigned* minmax, void* stream);
/* coarse weights -> maxpool/avgpool smoothing -> inverse-CDF fine depths (renderer.py:249-308).  u [n_rays][S_imp]. */
int b200_ray_importance(const float* t_c, const float* sigma_c, const float* u, float* t_f, long n_rays, int S, int S_imp,
                        void* stream);
/* merge coarse+fine by depth (renderer.py:212-222) and composite (ray_marcher.py:25-57): feat [n_rays][32], depth, wsum. */
int b200_ray_composite_fwd(const float* t_c, const float* sigma_c, const float* rgb_c, int S1, const float* t_f,
                           const float* sigma_f, const float* rgb_f, int S2, const unsigned* minmax, int white_back,
                           long n_rays, float* feat, float* depth, float* wsum, void* stream);
int b200_ray_composite_bwd(const float* t_c, const float* sigma_c, const float* rgb_c, int S1, const float* t_f,
                           const float* sigma_f, const float* rgb_f, int S2, const unsigned* minmax, int white_back,
                           long n_rays, const float* d_feat, const float* d_depth, const float* d_wsum, float* d_rgb_c,
                           float* d_sigma_c, float* d_rgb_f, float* d_sigma_f, void* stream);

/* ---- caller-side losses: training/coaches/base_coach.py:101-126 calc_loss (without LPIPS), :294-305 tv ---- */
/* loss[4] (ACCUMULATED; zero first) = {l2_lambda*(mse_image+mse_raw) + tv_lambda*tv, mse(image, real), mse(raw, area_R(real)), tv(depth)}.
 * image [n][C][H][W] and raw [n][C][R][R] are addressed through HOST arrays of 4 element strides {n, c, h, w} (NCHW views of NHWC
 * memory need no copy); depth [n][R][R] and real [n][C][H][W] contiguous.  image, raw and depth may each be NULL. */
int b200_pti_loss_fwd(const float* image, const long* image_strides, const float* raw, const long* raw_strides, const float* depth,
                      const float* real, int n, int C, int H, int W, int R, float l2_lambda, float tv_lambda, float* loss,
                      void* stream);
/* gradients of loss[0] scaled by the device scalar *dloss; written (not accumulated) with the given element strides. */
int b200_pti_loss_bwd(const float* image, const long* image_strides, const float* raw, const long* raw_strides, const float* depth,
                      const float* real, int n, int C, int H, int W, int R, float l2_lambda, float tv_lambda, const float* dloss,
                      float* d_image, const long* d_image_strides, float* d_raw, const long* d_raw_strides, float* d_depth,
                      void* stream);

/* ---- optimiser (training/coaches/base_coach.py:96-99, training/projectors/w_projector.py:134-140: torch.optim.Adam) ---------- */
/* One Adam update (torch.optim.Adam rule, amsgrad off) of `count` tensors in one launch (one per 320 tensors).  tensors: HOST array
 * of records of device pointers; records with n == 0 are skipped.  lr_dev: device scalar learning rate, or NULL to use `lr`.
 * step: device float = updates applied so far (0 at the start), advanced by one per call on the device, so a captured CUDA graph
 * replays the same launch every step; ticket: device uint32, zero-initialised, owned by the optimiser. */
typedef struct { float* p; const float* g; float* m; float* v; long n; } B200AdamTensor;
int b200_adam_step(const B200AdamTensor* tensors, int count, const float* lr_dev, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float* step, unsigned* ticket, void* stream);

/* ---- stage-1 (w-projection) caller-side kernels ------------------------------------------------- */
/* training/warping_loss.py:18-43 fused: rays of the predicted camera `ext` (ray_sampler.py:24-73), surface point o + d*depth
 * (:75-93), line-plane intersection with the canonical image plane (warping_loss.py:58-72), projection through w2c =
 * inverse(init_ext) and the intrinsics K -> uv [R*R][2] in [-1, 1].  ext / init_ext / w2c: device [16]; K: device [9]; depth:
 * device [R*R].  min_ndotu (device scalar, initialise to a large float; may be NULL) receives min |n.v| for the reference's
 * "no intersection" check. */
int b200_warp_uv_fwd(const float* ext, const float* init_ext, const float* w2c, const float* K, const float* depth, int R,
                     float* uv, float* min_ndotu, void* stream);
/* d_ext [16] ACCUMULATED (zero first), d_depth [R*R] written. */
int b200_warp_uv_bwd(const float* ext, const float* init_ext, const float* w2c, const float* K, const float* depth, int R,
                     const float* d_uv, float* d_ext, float* d_depth, void* stream);
/* Noise regulariser of w_projector.py:221-237 over nbuf square buffers (side sizes[b], powers of two): for every level of the 2x2
 * average-pool pyramid down to 8x8, (mean(n*roll(n,1,x)))^2 + (mean(n*roll(n,1,y)))^2, summed into *reg (ACCUMULATED).
 * bufs / dbufs: HOST arrays of device pointers; sizes: HOST array; sizes_dev: the same on the device; work / gwork: device
 * workspaces of b200_noise_pyramid_work_floats(nbuf, sizes) floats; sums: device [nbuf][8][2], ZEROED by the caller and
 * passed unchanged to the backward, which writes (*dreg) * d reg / d buffer into dbufs. */
long b200_noise_pyramid_work_floats(int nbuf, const int* sizes);
int b200_noise_pyramid_fwd(int nbuf, const float* const* bufs, const int* sizes, const int* sizes_dev, float* work, float* sums,
                           float* reg, void* stream);
int b200_noise_pyramid_bwd(int nbuf, const float* const* bufs, const int* sizes, const float* work, const float* sums,
                           const float* dreg, float* const* dbufs, float* gwork, void* stream);
/* w_projector.py:262-268 in place: buf <- (buf - mean) * rsqrt(mean((buf - mean)^2)); stats: device [nbuf][2], ZEROED by the caller. */
int b200_noise_normalize(int nbuf, float* const* bufs, const int* sizes, float* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200EG3D_H */
