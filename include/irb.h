/*
 * irb.h -- C ABI of the B200-native render-layer + bilateral-solver library
 * (libirb_b200.so, built from inverserenderingofindoorscene_b200/csrc/).
 *
 * This is the drop-in boundary for the hot path of lzqsd/InverseRenderingOfIndoorScene
 * (SURVEY.md 8b).  The reference is pure Python: the "FFI" a maintainer binds is ctypes
 * (see INTEGRATION.md).  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers to contiguous fp32 (unless stated), NCHW-style
 *     layouts exactly as the reference's PyTorch tensors; the caller owns every buffer
 *     (outputs and workspaces included) -- the library never allocates device memory;
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched asynchronously on it,
 *     no entry point synchronises the device except where stated;
 *   - the device is the caller's current device (cudaSetDevice done by the caller);
 *   - return 0 on success, negative for argument errors (IRB_E_*), positive = cudaError_t;
 *     a thread-local message is available from irb_last_error();
 *   - gradient outputs may be NULL = "not needed"; they are OVERWRITTEN, not accumulated;
 *   - env-map tensors (env, g_env, env_gt: [B,3,R,C,envH,envW]) must be 16-byte aligned (they are accessed as 16-byte
 *     vectors; IRB_E_ARG otherwise) -- every torch allocation is; the other tensors have no alignment requirement (a
 *     misaligned plane only disables the TMA-staged fast path);
 *   - B batch, K SG lobes, R x C render (env) grid, envH x envW hemisphere table
 *     (D = envH*envW <= IRB_MAX_DIRS), direction index d = i*envW + j.
 */
#ifndef IRB_H_
#define IRB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRB_ABI_VERSION 2
#define IRB_MAX_DIRS 128
#define IRB_MAX_LOBES 48

#define IRB_E_ARG (-1)     /* bad shape / null pointer */
#define IRB_E_UNSUPPORTED (-2)
#define IRB_E_WORKSPACE (-3)

int irb_abi_version(void);
const char* irb_last_error(void);
/* number of CUDA kernels this library has launched in this process so far (all threads); bench.py reports the
 * difference over its timed region as `gpu_launches`. */
long long irb_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * SG lobes -> hemisphere env-map.  Replaces models.output2env.fromSGtoIm (models.py:371-389)
 * and, with warp=1, models.output2env.output2env (models.py:391-404):
 *   lam = tan(pi/2*0.999*lamb), W = tan(pi/2*0.999*weight)   (warp=1; identity if warp=0)
 *   env[b,c,r,x,d] = sum_k W[b,3k+c,r,x] * exp(lam[b,k,r,x]*(axis[b,k,:,r,x].ls[d] - 1))
 * axis [B,K,3,R,C], lamb [B,K,R,C], weight [B,3K,R,C], env [B,3,R,C,envH,envW].
 * lamb_w / weight_w (optional, may be NULL): the warped lam / W ([B,K,R,C] / [B,3K,R,C]) that
 * output2env returns as its 3rd/4th results.
 * ------------------------------------------------------------------------------------- */
int irb_sg_to_env_fwd(const float* axis, const float* lamb, const float* weight, float* env,
                      float* lamb_w, float* weight_w,
                      int B, int K, int R, int C, int envH, int envW, int warp, void* stream);

/* Backward of the above (autograd of models.py:371-404).  g_env [B,3,R,C,envH,envW];
 * g_axis / g_lamb / g_weight shaped like the inputs; gradients are w.r.t. the RAW inputs
 * when warp=1 (chain through tan), w.r.t. lam / W when warp=0. */
int irb_sg_to_env_bwd(const float* axis, const float* lamb, const float* weight, const float* g_env,
                      float* g_axis, float* g_lamb, float* g_weight,
                      int B, int K, int R, int C, int envH, int envW, int warp, void* stream);

/* ---------------------------------------------------------------------------------------
 * Microfacet shading of a materialised env-map.  Replaces the body of
 * models.renderingLayer.forwardEnv after pooling (models.py:466-520).
 * albedo/normal [B,3,R,C], rough [B,1,R,C] are the POOLED maps (see irb_avgpool_*);
 * env [B,3,R,C,envH,envW]; view [3,R,C] = renderingLayer.v (models.py:428-432);
 * out_diff / out_spec [B,3,R,C].
 * ------------------------------------------------------------------------------------- */
int irb_render_env_fwd(const float* albedo, const float* normal, const float* rough, const float* env,
                       const float* view, float* out_diff, float* out_spec,
                       int B, int R, int C, int envH, int envW, float F0, void* stream);

int irb_render_env_bwd(const float* albedo, const float* normal, const float* rough, const float* env,
                       const float* view, const float* g_diff, const float* g_spec,
                       float* g_albedo, float* g_normal, float* g_rough, float* g_env,
                       int B, int R, int C, int envH, int envW, float F0, void* stream);

/* ---------------------------------------------------------------------------------------
 * FUSED SG-lit render (the north-star kernel): output2env.output2env followed by
 * renderingLayer.forwardEnv (wrapperBRDFLight.py:177,194) without materialising the env-map.
 * Inputs as above; lamb / weight are the RAW decoder outputs when warp=1.
 * lobe_uv (optional, may be NULL in both calls): [B,2K,R,C] side tensor the forward fills with the per-lobe hemisphere
 * sums U_k = sum_d w_d e_kd (plane 2k) and V_k = sum_d w_d spec_d e_kd (plane 2k+1); a backward that is handed the same
 * tensor skips rebuilding them (ABI 2; 10 % less FMA-pipe work in the backward, 96 B/pixel of extra traffic at K = 12).
 * ------------------------------------------------------------------------------------- */
int irb_render_sg_fwd(const float* albedo, const float* normal, const float* rough,
                      const float* axis, const float* lamb, const float* weight, const float* view,
                      float* out_diff, float* out_spec, float* lobe_uv,
                      int B, int K, int R, int C, int envH, int envW, float F0, int warp, void* stream);

int irb_render_sg_bwd(const float* albedo, const float* normal, const float* rough,
                      const float* axis, const float* lamb, const float* weight, const float* view,
                      const float* g_diff, const float* g_spec, const float* lobe_uv,
                      float* g_albedo, float* g_normal, float* g_rough,
                      float* g_axis, float* g_lamb, float* g_weight,
                      int B, int K, int R, int C, int envH, int envW, float F0, int warp, void* stream);

/* ---------------------------------------------------------------------------------------
 * F.adaptive_avg_pool2d as used at models.py:465-469 (window [floor(i*H/R), ceil((i+1)*H/R))).
 * in [N,H,W] -> out [N,R,C] (N = B*channels); bwd spreads g_out uniformly over each window.
 * ------------------------------------------------------------------------------------- */
int irb_avgpool_fwd(const float* in, float* out, int N, int H, int W, int R, int C, void* stream);
int irb_avgpool_bwd(const float* g_out, float* g_in, int N, int H, int W, int R, int C, void* stream);
/* the three maps forwardEnv pools (albedo, normal, rough: models.py:465-469) in ONE launch; tensor k has n_k = B*channels
 * planes; a NULL in/out pair is skipped.  bwd: g* at R x C -> gi* at H x W. */
int irb_avgpool3_fwd(const float* in0, const float* in1, const float* in2, float* out0, float* out1, float* out2, int n0, int n1,
                     int n2, int H, int W, int R, int C, void* stream);
int irb_avgpool3_bwd(const float* g0, const float* g1, const float* g2, float* gi0, float* gi1, float* gi2, int n0, int n1,
                     int n2, int H, int W, int R, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused scale-ambiguity regression.  Replaces models.LSregressDiffSpec (models.py:23-84) for the
 * reference's call pattern (wrapperBRDFLight.py:197-201: the statistics come from detached tensors):
 * mask = im < 0.9; per-sample 2x2 normal equations on (diff, spec) -> (cD, cS) with the
 * diffuse-only fallback and clamps of models.py:44-64; second scalar regression cI of
 * clamp(cD diff + cS spec, 0, 1) onto im (models.py:70-78);
 * out_diff = cI cD diff_orig, out_spec = cI cS spec_orig; coef[B,2] = (cI cD, cI cS) for the backward.
 * All image pointers [B, n_per_sample]; stats = caller scratch of 8*B doubles.
 * ------------------------------------------------------------------------------------- */
int irb_lsregress_diffspec(const float* diff, const float* spec, const float* im, const float* diff_orig,
                           const float* spec_orig, float* out_diff, float* out_spec, float* coef, double* stats,
                           int B, long long n_per_sample, void* stream);

/* Fused render loss of the light branch.  Replaces, with three launches and no host sync,
 *   dS, sS = models.LSregressDiffSpec(diff.detach(), spec.detach(), im, diff, spec)      (wrapperBRDFLight.py:197-201)
 *   rend   = clamp(dS + sS, 0, 1)                                                          (wrapperBRDFLight.py:203)
 *   num    = sum((rend - im)^2 * seg),  cnt = sum(seg)     (wrapperBRDFLight.py:192,205-207: renderErr = num/cnt/3,
 *            the reference fetches cnt = pixelNum to the host with .item(); here it stays on the device)
 * diff / spec / im [B,ch,n_pix], seg [B,n_pix] (broadcast over channels).  Outputs: rendered [B,ch,n_pix] (NULL = skip),
 * gbase [B,ch,n_pix] = d num / d(dS+sS) (NULL = skip), coef [B,2] = (cI cD, cI cS), terms[2] fp64 = {num, cnt} (device),
 * stats = scratch of 8*B doubles.
 * irb_render_loss_bwd: g_diff = *g_num * coef[b,0] * gbase, g_spec = *g_num * coef[b,1] * gbase (g_num = device scalar
 * dL/dnum, NULL = 1; the coefficients are graph constants as in models.py:58,76). */
int irb_render_loss_fwd(const float* diff, const float* spec, const float* im, const float* seg, float* rendered, float* gbase,
                        float* coef, double* stats, double* terms, int B, int ch, long long n_pix, void* stream);
int irb_render_loss_bwd(const float* gbase, const float* coef, const float* g_num, float* g_diff, float* g_spec, int B, int ch,
                        long long n_pix, void* stream);

/* models.LSregress (models.py:7-21): coef[b] = clamp(sum(pred*gt)/clamp(sum(pred^2),1e-5), 1e-3, 1e3) per sample,
 * out = origin * coef (coef is a constant of the autograd graph, models.py:13 .detach()).  pred / gt / origin / out
 * [B, n_per_sample]; coef [B]; stats = caller scratch of 2*B doubles. */
int irb_lsregress(const float* pred, const float* gt, const float* origin, float* out, float* coef, double* stats,
                  int B, long long n_per_sample, void* stream);

/* Fused reconstruction loss of the light branch: replaces, for an env-map that comes from SG lobes, the chain
 *   envPred = output2env(axis, lamb, weight)                                   (wrapperBRDFLight.py:172)
 *   envScaled = LSregress(envPred.detach()*seg, envGT*seg, envPred)            (wrapperBRDFLight.py:179-181)
 *   num = sum((log(envScaled + off) - log(envGT + off))^2 * seg)               (wrapperBRDFLight.py:183-187)
 * without materialising envPred.  env_gt [B,3,R,C,envH,envW]; seg [B,1,R,C] (mask or fractional weight);
 * stats [B,4] fp64 out = {sum(E G s^2), sum(E^2 s^2), num_b, sum(s)}; coef_b = clamp(stats0/max(stats1,1e-5), 1e-3, 1e3).
 * fwd: when any of g_axis / g_lamb / g_weight is non-null the second pass also writes d(sum_b num_b)/d(parameter)
 * (unit upstream gradient; the caller scales these small tensors), so a training step needs two passes, not three.
 * bwd: g_num = device scalar dL/d(sum_b num_b); writes the gradients w.r.t. the ORIGINAL SG parameters (coef is a
 * graph constant as in models.py:13).  Needs K <= 24, R*C % 4 == 0, 16-byte aligned buffers (else IRB_E_UNSUPPORTED).
 * Pixels with seg == 0 contribute exactly zero and their env_gt values are never looked at (a NaN / Inf there stays out of
 * the loss and of the gradients; the reference's `(...)**2 * seg` would propagate it). */
int irb_sg_reconst_fwd(const float* axis, const float* lamb, const float* weight, const float* env_gt, const float* seg,
                       double* stats, float* g_axis, float* g_lamb, float* g_weight, int B, int K, int R, int C, int envH,
                       int envW, int warp, float offset, void* stream);
int irb_sg_reconst_bwd(const float* axis, const float* lamb, const float* weight, const float* env_gt, const float* seg,
                       const double* stats, const float* g_num, float* g_axis, float* g_lamb, float* g_weight, int B, int K,
                       int R, int C, int envH, int envW, int warp, float offset, void* stream);

/* SG lighting -> diffuse irradiance ("shading") image: replaces utils.predToShading (utils.py:156-195, called from
 * testReal.py:638).  pred [B,7K,R,C] = cat(axis[K,3], lambOrig[K], weightOrig[K,3]) per image (the reference passes
 * B = 1); out [B,3,R,C] = max(sum_d cosEl_d sinEl_d * env_d, 0) over an envH x envW hemisphere table (16 x 32 in the
 * reference; envH*envW <= 2048).  Forward only. */
int irb_sg_shading(const float* pred, float* out, int B, int K, int R, int C, int envH, int envW, void* stream);

/* ---------------------------------------------------------------------------------------
 * Bilateral-grid solver.  Replaces BilateralGrid.BilateralGrid / bistochastize /
 * BilateralSolver.solve / solveGrad (BilateralGrid.py:44-207) and the per-image bridge
 * BilateralLayer.BilateralFunction.forward/backward (BilateralLayer.py:24-124).
 *
 * irb_bg_workspace_bytes: size of the opaque grid workspace for B images of H x W.
 * irb_bg_build: guide [B,3,H,W]; the grid image is fl32(guide * guide_scale) -- pass 255.0f for the
 *   BilateralLayer call (BilateralLayer.py:52 multiplies the [0,1] guide by 255.0 in float32), 1.0f for an
 *   image that is already in [0,255]; builds per image the 5-D XYLUV grid (fp64 coordinates, BilateralGrid.py:44-59),
 *   vertex ids = rank of the hash among sorted unique hashes (BilateralGrid.py:61-65),
 *   the +-1 neighbour table (:72-84) and the bistochastisation vectors (:109-120).
 *   nverts_out: HOST int[B] or NULL; when given it is written after a stream synchronise (the only sync).
 * irb_bg_vertex_index: copies the per-pixel vertex ids (int32 [B,H*W]) for inspection.
 * irb_bg_solve_fwd: pred [B,ch,H,W], conf [B,1,H,W] -> out [B,ch,H,W];
 *   yhat (fp64, [B, ch, H*W] capacity) receives the bilateral-space solution saved for backward.
 * irb_bg_solve_bwd: adjoint solve (BilateralGrid.py:152-184) -> g_pred [B,ch,H,W], g_conf [B,1,H,W].
 * cg_maxiter is an explicit integer (the reference reads it from the wrong slot,
 * BilateralLayer.py:44,79).
 * ------------------------------------------------------------------------------------- */
size_t irb_bg_workspace_bytes(int B, int H, int W, int ch);
int irb_bg_build(const float* guide, float guide_scale, int B, int H, int W, double sigma_luma, double sigma_chroma,
                 double sigma_spatial, void* ws, size_t ws_bytes, int* nverts_out, void* stream);
int irb_bg_vertex_index(const void* ws, int32_t* idx_out, int B, int H, int W, void* stream);
int irb_bg_solve_fwd(void* ws, const float* pred, const float* conf, int B, int ch, int H, int W,
                     double lam, double A_diag_min, double cg_tol, int cg_maxiter,
                     float* out, double* yhat, void* stream);
int irb_bg_solve_bwd(void* ws, const float* pred, const float* conf, const float* g_out, const double* yhat,
                     int B, int ch, int H, int W, double lam, double A_diag_min, double cg_tol, int cg_maxiter,
                     float* g_pred, float* g_conf, void* stream);

/* Primitive operators of a built grid (BilateralGrid.py:90-107), fp64 device buffers [B,ch,H*W] (vertex vectors use the
 * first V entries of each plane, the rest is zero): op 0 = splat (S x, pixel -> vertex), 1 = slice (S^T y, vertex -> pixel),
 * 2 = blur (2 dim y + sum of the five [1 0 1] neighbour sums, vertex -> vertex).  x and out must not alias. */
int irb_bg_apply(const void* ws, int op, const double* x, double* out, int B, int ch, int H, int W, void* stream);
/* the bistochastisation vectors n, m of BilateralGrid.bistochastize (BilateralGrid.py:109-120), device fp64 [B,H*W] */
int irb_bg_bistochastization(const void* ws, double* n, double* m, int B, int H, int W, void* stream);
/* per image: number of vertices V (device int[B]) and of directed neighbour links E (device uint64[B]) -- SURVEY 8d */
int irb_bg_stats(const void* ws, int* nverts, unsigned long long* nlinks, int B, int H, int W, void* stream);

/* per image (device int[B]): 1 if some pixel's grid coordinate left [0, 255) when the grid was built -- i.e. guide *
 * guide_scale outside [0, 255] (or sigma too small for the image size).  The reference's hash (BilateralGrid.py:57,86-88)
 * aliases there; this library then gives unspecified vertex ids.  The build itself never synchronises to report it. */
int irb_bg_range_flags(const void* ws, int* flags, int B, int H, int W, void* stream);
/* test / debugging aid: PCG scalars {rho, <p,q>, <r,r>, -} per (image, channel, iteration) of the last solve on this
 * workspace, device fp64 [B][4][258][4]; slot 257 = {<b,b>, -, -, -} */
int irb_bg_debug_scalars(const void* ws, double* out, int B, int ch, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IRB_H_ */
