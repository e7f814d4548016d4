// irb_render_sg.cu -- FUSED SG-lit microfacet render, forward and hand-written backward
// (replaces models.output2env.output2env + models.renderingLayer.forwardEnv,
// models.py:371-404 and 466-520, without materialising the [B,3,R,C,8,16] env-map).
//
// Mapping: one thread per render-grid pixel, 128-thread CTAs, grid = (ceil(R*C/128), B).
// All tensors are channel-planar [B,ch,R,C], so the warp's 32 pixels read/write 128 contiguous
// bytes per plane (fully coalesced, each byte of HBM traffic touched exactly once).  The lobe
// parameters, their gradient accumulators and the BRDF state are register-resident for the whole
// 128-direction sweep; the hemisphere table lives in the kernel-parameter constant bank.
// The kernel is FP32-FMA / MUFU bound: ~100 FLOP per HBM byte.
#include "irb_common.cuh"
#include "irb_driver.cuh"

namespace irb {

constexpr int kThreads = 128;

template <int KC>
__global__ void __launch_bounds__(kThreads) render_sg_fwd_kernel(const RenderSgArgs a, const __grid_constant__ EnvTab tab) {
    const size_t pix = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (pix >= (size_t)a.R * a.C) return;
    drv_render_sg_fwd<float, KC>(a, tab, blockIdx.y, pix);
}

template <int KP, int KL, bool PIX, bool LOBE>
__global__ void __launch_bounds__(kThreads) render_sg_bwd_kernel(const RenderSgArgs a, const __grid_constant__ EnvTab tab) {
    const size_t pix = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (pix >= (size_t)a.R * a.C) return;
    drv_render_sg_bwd<float, KP, KL, PIX, LOBE>(a, tab, blockIdx.y, pix);
}

static int check_common(const RenderSgArgs& a, int envH, int envW, EnvTab& tab) {
    IRB_REQUIRE(a.B > 0 && a.K > 0 && a.R > 0 && a.C > 0, "render_sg: non-positive shape B=%d K=%d R=%d C=%d", a.B, a.K, a.R, a.C);
    IRB_REQUIRE(a.albedo && a.normal && a.rough && a.axis && a.lamb && a.weight && a.view, "render_sg: null input pointer");
    if (a.K > 24) {
        set_error("render_sg: fused path holds at most 24 lobes in registers (K=%d); use sg_to_env + render_env", a.K);
        return IRB_E_UNSUPPORTED;
    }
    if (a.B > 65535) {
        set_error("render_sg: B=%d exceeds grid.y", a.B);
        return IRB_E_UNSUPPORTED;
    }
    if (!make_env_tab(tab, envH, envW)) {
        set_error("render_sg: env table %dx%d unsupported (need envH*envW <= %d, multiple of 4)", envH, envW, IRB_MAX_DIRS);
        return IRB_E_UNSUPPORTED;
    }
    return 0;
}

}  // namespace irb

using namespace irb;

extern "C" int irb_render_sg_fwd(const float* albedo, const float* normal, const float* rough, const float* axis,
                                 const float* lamb, const float* weight, const float* view, float* out_diff,
                                 float* out_spec, int B, int K, int R, int C, int envH, int envW, float F0, int warp,
                                 void* stream) {
    RenderSgArgs a{};
    a.albedo = albedo; a.normal = normal; a.rough = rough; a.axis = axis; a.lamb = lamb; a.weight = weight; a.view = view;
    a.out_diff = out_diff; a.out_spec = out_spec;
    a.B = B; a.K = K; a.R = R; a.C = C; a.F0 = F0; a.warp = warp;
    EnvTab tab;
    if (int rc = check_common(a, envH, envW, tab)) return rc;
    IRB_REQUIRE(out_diff && out_spec, "render_sg_fwd: null output pointer");
    const dim3 grid((unsigned)(((size_t)R * C + kThreads - 1) / kThreads), (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 12) render_sg_fwd_kernel<12><<<grid, kThreads, 0, st>>>(a, tab);
    else render_sg_fwd_kernel<24><<<grid, kThreads, 0, st>>>(a, tab);
    return check_launch("render_sg_fwd");
}

extern "C" int irb_render_sg_bwd(const float* albedo, const float* normal, const float* rough, const float* axis,
                                 const float* lamb, const float* weight, const float* view, const float* g_diff,
                                 const float* g_spec, float* g_albedo, float* g_normal, float* g_rough, float* g_axis,
                                 float* g_lamb, float* g_weight, int B, int K, int R, int C, int envH, int envW, float F0,
                                 int warp, void* stream) {
    RenderSgArgs a{};
    a.albedo = albedo; a.normal = normal; a.rough = rough; a.axis = axis; a.lamb = lamb; a.weight = weight; a.view = view;
    a.g_diff = g_diff; a.g_spec = g_spec;
    a.g_albedo = g_albedo; a.g_normal = g_normal; a.g_rough = g_rough; a.g_axis = g_axis; a.g_lamb = g_lamb; a.g_weight = g_weight;
    a.B = B; a.K = K; a.R = R; a.C = C; a.F0 = F0; a.warp = warp;
    EnvTab tab;
    if (int rc = check_common(a, envH, envW, tab)) return rc;
    IRB_REQUIRE(g_diff && g_spec, "render_sg_bwd: null upstream gradient");
    const bool pix = g_normal || g_rough;
    const bool lobe = g_axis || g_lamb || g_weight || g_albedo;
    if (!pix && !lobe) return 0;
    const dim3 grid((unsigned)(((size_t)R * C + kThreads - 1) / kThreads), (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 12) {
        if (pix && lobe) render_sg_bwd_kernel<12, 12, true, true><<<grid, kThreads, 0, st>>>(a, tab);
        else if (pix) render_sg_bwd_kernel<12, 12, true, false><<<grid, kThreads, 0, st>>>(a, tab);
        else render_sg_bwd_kernel<12, 12, false, true><<<grid, kThreads, 0, st>>>(a, tab);
    } else {
        if (pix && lobe) render_sg_bwd_kernel<24, 12, true, true><<<grid, kThreads, 0, st>>>(a, tab);
        else if (pix) render_sg_bwd_kernel<24, 12, true, false><<<grid, kThreads, 0, st>>>(a, tab);
        else render_sg_bwd_kernel<24, 12, false, true><<<grid, kThreads, 0, st>>>(a, tab);
    }
    return check_launch("render_sg_bwd");
}
