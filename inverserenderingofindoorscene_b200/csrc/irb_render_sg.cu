// irb_render_sg.cu -- FUSED SG-lit microfacet render, forward and hand-written backward
// (replaces models.output2env.output2env + models.renderingLayer.forwardEnv,
// models.py:371-404 and 466-520, without materialising the [B,3,R,C,8,16] env-map).
//
// Mapping: one thread per render-grid pixel; all tensors are channel-planar [B,ch,R,C], so a tile of
// 128 consecutive pixels is 128 contiguous floats (512 B) in every plane.
//
// Primary path: PERSISTENT, TMA-STAGED kernels.  grid = (#SM x resident CTAs); each 128-thread CTA
// walks tiles of 128 pixels.  The 16+7K input planes of a tile (BRDF maps, view vectors, SG lobes,
// upstream gradients) are staged into shared memory by bulk async copies (cp.async.bulk ->
// UBLKCP, completion on an mbarrier) into a 2-stage ring: the copies of tile t+1 are issued before
// the direction sweep of tile t starts, so the ~100 plane reads a pixel needs never sit on the critical
// path (the non-staged kernel spent ~45 % of its time in long-scoreboard stalls on exactly those reads
// with the 2 warps/scheduler that 250 registers/thread allow; profiles/r01).  Each byte is read from
// HBM exactly once; results go back with coalesced 128-byte stores straight from registers.
// The sweep itself (lobes, gradient accumulators, BRDF state) is register-resident; the hemisphere
// table lives in the kernel-parameter constant bank.  FP32-FMA / MUFU bound.
//
// Fallback path (same math, same results): plain one-thread-per-pixel kernels with direct global
// loads, used when a plane is not 16-byte addressable (R*C % 4 != 0 or misaligned base pointers).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "irb_common.cuh"
#include "irb_driver.cuh"
#include "irb_tma.cuh"
#include "irb_launch.cuh"

// Resident CTAs per SM the forward / the lobe-only backward are compiled for.  Measured on one B200 (profiles/r02): forward
// 4 / 5 / 6 -> 0.854 / 0.832 / 0.903 ms (5 = 96 registers; shared memory still caps residency at 4 CTAs, the tighter
// register budget itself schedules better); lobe-only backward 2 / 3 / 4 -> 0.615 / 0.611 / 0.648 ms at 16 x 120 x 160.
#ifndef IRB_FWD_MINB
#define IRB_FWD_MINB 5
#endif
#ifndef IRB_BWD_MINB
#define IRB_BWD_MINB 2
#endif
#ifndef IRB_LOBEONLY_MINB
#define IRB_LOBEONLY_MINB 3
#endif

namespace irb {

constexpr int kThreads = 128;
constexpr int kTile = 128;  // pixels per tile = threads per CTA

// ---------------------------------------------------------------------------------------------
// plain kernels (fallback)
// ---------------------------------------------------------------------------------------------
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

// plane order inside a stage (each plane = kTile floats): albedo 3 | normal 3 | rough 1 | view 3 |
// axis 3K | lamb K | weight 3K | g_diff 3 | g_spec 3 (backward only)
__host__ __device__ inline int stage_planes(int K, bool bwd) { return 10 + 7 * K + (bwd ? 6 : 0); }

// Issue the bulk copies of one tile (executed by warp 0; lanes split the planes).
__device__ __forceinline__ void issue_tile(float* stage, uint64_t* bar, const RenderSgArgs& a, size_t tile, size_t tiles_per_img,
                                           bool bwd) {
    const int lane = threadIdx.x & 31;
    const size_t plane = (size_t)a.R * a.C;
    const size_t b = tile / tiles_per_img, pix0 = (tile % tiles_per_img) * kTile;
    const uint32_t valid = (uint32_t)((plane - pix0) < (size_t)kTile ? (plane - pix0) : (size_t)kTile);
    const uint32_t bytes = valid * 4u;  // multiple of 16: the launcher requires plane % 4 == 0
    const int K = a.K;
    const int np = stage_planes(K, bwd);
    if (lane == 0) mbar_expect_tx(bar, bytes * (uint32_t)np);
    __syncwarp();
    for (int p = lane; p < np; p += 32) {
        const float* src;
        if (p < 3) src = a.albedo + (b * 3 + p) * plane;
        else if (p < 6) src = a.normal + (b * 3 + (p - 3)) * plane;
        else if (p < 7) src = a.rough + b * plane;
        else if (p < 10) src = a.view + (size_t)(p - 7) * plane;
        else if (p < 10 + 3 * K) src = a.axis + (b * 3 * K + (p - 10)) * plane;
        else if (p < 10 + 4 * K) src = a.lamb + (b * K + (p - 10 - 3 * K)) * plane;
        else if (p < 10 + 7 * K) src = a.weight + (b * 3 * K + (p - 10 - 4 * K)) * plane;
        else if (p < 13 + 7 * K) src = a.g_diff + (b * 3 + (p - 10 - 7 * K)) * plane;
        else src = a.g_spec + (b * 3 + (p - 13 - 7 * K)) * plane;
        bulk_g2s(stage + (size_t)p * kTile, src + pix0, bytes, bar);
    }
}

// input view of a stage: same field meaning as the global tensors, with b = 0, plane = kTile
__device__ __forceinline__ RenderSgArgs stage_view(const RenderSgArgs& a, float* stage) {
    RenderSgArgs s = a;
    const int K = a.K;
    s.albedo = stage;
    s.normal = stage + 3 * kTile;
    s.rough = stage + 6 * kTile;
    s.view = stage + 7 * kTile;
    s.axis = stage + 10 * kTile;
    s.lamb = stage + (10 + 3 * K) * kTile;
    s.weight = stage + (10 + 4 * K) * kTile;
    s.g_diff = stage + (10 + 7 * K) * kTile;
    s.g_spec = stage + (13 + 7 * K) * kTile;
    return s;
}

}  // namespace irb
#include "irb_render_sg_group.cuh"
namespace irb {

// MODE 0: forward<KP>; MODE 1: backward<KP,KL,PIX,LOBE>   (single-thread-per-pixel staged variant; kept for A/B via
// IRB_RENDER_VARIANT=tma1)
template <int MODE, int KP, int KL, bool PIX, bool LOBE>
__global__ void __launch_bounds__(kThreads) render_sg_tma_kernel(const RenderSgArgs a, const __grid_constant__ EnvTab tab,
                                                                  const unsigned total_tiles, const unsigned tiles_per_img) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2];
    constexpr bool BWD = (MODE == 1);
    const int np = stage_planes(a.K, BWD);
    float* stage[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw) + (size_t)np * kTile};
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t plane = (size_t)a.R * a.C;
    unsigned t = blockIdx.x;
    if (t < total_tiles && threadIdx.x < 32) issue_tile(stage[0], &full_bar[0], a, t, tiles_per_img, BWD);
    for (unsigned it = 0; t < total_tiles; ++it, t += gridDim.x) {
        const int s = it & 1;
        const unsigned tn = t + gridDim.x;
        if (tn < total_tiles && threadIdx.x < 32) {
            // stage s^1 was last read (generic proxy) in iteration it-1, which ended with __syncthreads()
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile(stage[s ^ 1], &full_bar[s ^ 1], a, tn, tiles_per_img, BWD);
        }
        mbar_wait(&full_bar[s], (it >> 1) & 1);
        const size_t b = t / tiles_per_img, pix0 = (size_t)(t % tiles_per_img) * kTile;
        if (pix0 + threadIdx.x < plane) {
            const RenderSgArgs sv = stage_view(a, stage[s]);
            const PixView vi{0, threadIdx.x, (size_t)kTile};
            const PixView vo{b, pix0 + threadIdx.x, plane};
            if (MODE == 0) drv_render_sg_fwd<float, KP>(sv, vi, a, vo, tab);
            else drv_render_sg_bwd<float, KP, KL, PIX, LOBE>(sv, vi, a, vo, tab);
        }
        __syncthreads();  // everyone is done with stage s before it is refilled (iteration it+1 issues into it)
    }
}

struct Launch {
    int sms = 0;
};

// test knobs (environment): IRB_DISABLE_TMA=1 forces the plain kernels; IRB_TMA_MAX_CTAS=n caps the persistent
// grid so that small test problems still walk many tiles per CTA (exercises the 2-stage ring)
static bool tma_ok(const RenderSgArgs& a, bool bwd) {
    const char* dis = getenv("IRB_DISABLE_TMA");
    if (dis && dis[0] == '1') return false;
    const size_t plane = (size_t)a.R * a.C;
    if (plane % 4 != 0) return false;
    const void* ptrs[] = {a.albedo, a.normal, a.rough, a.view, a.axis, a.lamb, a.weight, bwd ? a.g_diff : a.albedo,
                          bwd ? a.g_spec : a.albedo};
    for (const void* p : ptrs)
        if (((uintptr_t)p & 15u) != 0) return false;
    return true;
}

template <typename KernelT>
static int launch_tma(KernelT kernel, const RenderSgArgs& a, const EnvTab& tab, bool bwd, cudaStream_t st, const char* what,
                      int tile = kTile) {
    const size_t plane = (size_t)a.R * a.C;
    const size_t tiles_per_img = (plane + tile - 1) / tile;
    const size_t total = tiles_per_img * (size_t)a.B;
    if (total >= 0xffffffffull) {
        set_error("%s: too many tiles", what);
        return IRB_E_UNSUPPORTED;
    }
    const size_t smem = 2 * (size_t)stage_planes(a.K, bwd) * tile * sizeof(float);
    int per_sm = 1;
    if (int e = persistent_ctas_per_sm(kernel, kThreads, smem, &per_sm)) {
        set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString((cudaError_t)e));
        return e;
    }
    size_t grid = (size_t)cached_sms() * per_sm;
    if (grid > total) grid = total;
    if (const char* cap = getenv("IRB_TMA_MAX_CTAS")) {
        const long c = atol(cap);
        if (c > 0 && (size_t)c < grid) grid = (size_t)c;
    }
    kernel<<<(unsigned)grid, kThreads, smem, st>>>(a, tab, (unsigned)total, (unsigned)tiles_per_img);
    return check_launch(what);
}

// lobe-split group kernels: same persistent launch, plus the tensor maps of the staged planes (IRB_TMA_TENSOR=0 keeps the
// per-plane bulk copies for A/B runs)
template <typename KernelT>
static int launch_group(KernelT kernel, const RenderSgArgs& a, const EnvTab& tab, bool bwd, cudaStream_t st, const char* what,
                        int tile, bool saved = false) {
    const size_t plane = (size_t)a.R * a.C;
    const size_t tiles_per_img = (plane + tile - 1) / tile;
    const size_t total = tiles_per_img * (size_t)a.B;
    if (total >= 0xffffffffull || plane >= 0x7fffffffull) {
        set_error("%s: too many tiles", what);
        return IRB_E_UNSUPPORTED;
    }
    TileMaps maps;
    memset(&maps, 0, sizeof(maps));
    const char* tt = getenv("IRB_TMA_TENSOR");
    if (!(tt && tt[0] == '0')) {
        const unsigned K = (unsigned)a.K, B = (unsigned)a.B, T = (unsigned)tile;
        bool ok = cached_tile_map(&maps.m[0], a.albedo, plane, 3, B, T) && cached_tile_map(&maps.m[1], a.normal, plane, 3, B, T) &&
                  cached_tile_map(&maps.m[2], a.rough, plane, 1, B, T) && cached_tile_map(&maps.m[3], a.view, plane, 3, 1, T) &&
                  cached_tile_map(&maps.m[4], a.axis, plane, 3 * K, B, T) && cached_tile_map(&maps.m[5], a.lamb, plane, K, B, T) &&
                  cached_tile_map(&maps.m[6], a.weight, plane, 3 * K, B, T);
        if (ok && bwd) ok = cached_tile_map(&maps.m[7], a.g_diff, plane, 3, B, T) && cached_tile_map(&maps.m[8], a.g_spec, plane, 3, B, T);
        if (ok && saved) ok = cached_tile_map(&maps.m[9], a.lobe_uv_in, plane, 2 * K, B, T);
        maps.on = ok ? 1 : 0;
    }
    const size_t smem = 2 * (size_t)(stage_planes(a.K, bwd) + (saved ? 2 * a.K : 0)) * tile * sizeof(float);
    int per_sm = 1;
    if (int e = persistent_ctas_per_sm(kernel, kThreads, smem, &per_sm)) {
        set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString((cudaError_t)e));
        return e;
    }
    size_t grid = (size_t)cached_sms() * per_sm;
    if (grid > total) grid = total;
    if (const char* cap = getenv("IRB_TMA_MAX_CTAS")) {
        const long c = atol(cap);
        if (c > 0 && (size_t)c < grid) grid = (size_t)c;
    }
    kernel<<<(unsigned)grid, kThreads, smem, st>>>(a, tab, maps, (unsigned)total, (unsigned)tiles_per_img);
    return check_launch(what);
}

// Kernel variant.  Measured on B200 (profiles/r01, config 2, fwd / bwd): lobe-split "group" kernels 0.82 / 1.88 ms,
// single-thread-per-pixel staged "tma1" 1.10 / 2.38 ms, plain (no staging) 1.23 / 3.41 ms -> group is the default.
// IRB_RENDER_VARIANT = group | tma1 | plain overrides it (A/B runs and tests).
//   returns 0 = group, 1 = tma1, 2 = plain
static int render_variant(bool bwd, int K) {
    (void)bwd;
    (void)K;
    const char* v = getenv("IRB_RENDER_VARIANT");
    if (v) {
        if (v[0] == 't') return 1;
        if (v[0] == 'p') return 2;
    }
    return 0;
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
                                 float* out_spec, float* lobe_uv, int B, int K, int R, int C, int envH, int envW, float F0,
                                 int warp, void* stream) {
    RenderSgArgs a{};
    a.albedo = albedo; a.normal = normal; a.rough = rough; a.axis = axis; a.lamb = lamb; a.weight = weight; a.view = view;
    a.out_diff = out_diff; a.out_spec = out_spec; a.lobe_uv_out = lobe_uv;
    a.B = B; a.K = K; a.R = R; a.C = C; a.F0 = F0; a.warp = warp;
    EnvTab tab;
    if (int rc = check_common(a, envH, envW, tab)) return rc;
    IRB_REQUIRE(out_diff && out_spec, "render_sg_fwd: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int variant = tma_ok(a, false) ? render_variant(false, K) : 2;
    if (variant == 0 && tab.D % (2 * (K <= 6 ? 1 : (K <= 12 ? 2 : 4))) != 0) variant = 1;  // group kernels walk 2*TPP directions per iteration
    if (variant == 1) {
        if (K <= 12) return launch_tma(render_sg_tma_kernel<0, 12, 12, false, false>, a, tab, false, st, "render_sg_fwd(tma)");
        return launch_tma(render_sg_tma_kernel<0, 24, 24, false, false>, a, tab, false, st, "render_sg_fwd(tma)");
    }
    if (variant == 0) {
        int rc;
        const char* w = "render_sg_fwd(group)";
#define IRB_GROUP_FWD(TPP)                                                                                                   \
    (K == TPP * kLPT ? launch_group(render_sg_group_kernel<0, TPP, false, false, IRB_FWD_MINB, false, true>, a, tab, false, st, w, kThreads / TPP) \
                     : launch_group(render_sg_group_kernel<0, TPP, false, false, IRB_FWD_MINB>, a, tab, false, st, w, kThreads / TPP))
        if (K <= 6) rc = IRB_GROUP_FWD(1);
        else if (K <= 12) rc = IRB_GROUP_FWD(2);
        else rc = IRB_GROUP_FWD(4);
#undef IRB_GROUP_FWD
        return rc;   // pixels outside the regular regime were shaded inside the kernel (group_general_pixel)
    }
    const dim3 grid((unsigned)(((size_t)R * C + kThreads - 1) / kThreads), (unsigned)B);
    if (K <= 12) render_sg_fwd_kernel<12><<<grid, kThreads, 0, st>>>(a, tab);
    else render_sg_fwd_kernel<24><<<grid, kThreads, 0, st>>>(a, tab);
    return check_launch("render_sg_fwd");
}

extern "C" int irb_render_sg_bwd(const float* albedo, const float* normal, const float* rough, const float* axis,
                                 const float* lamb, const float* weight, const float* view, const float* g_diff,
                                 const float* g_spec, const float* lobe_uv, float* g_albedo, float* g_normal, float* g_rough,
                                 float* g_axis, float* g_lamb, float* g_weight, int B, int K, int R, int C, int envH, int envW,
                                 float F0, int warp, void* stream) {
    RenderSgArgs a{};
    a.albedo = albedo; a.normal = normal; a.rough = rough; a.axis = axis; a.lamb = lamb; a.weight = weight; a.view = view;
    a.g_diff = g_diff; a.g_spec = g_spec; a.lobe_uv_in = lobe_uv;
    a.g_albedo = g_albedo; a.g_normal = g_normal; a.g_rough = g_rough; a.g_axis = g_axis; a.g_lamb = g_lamb; a.g_weight = g_weight;
    a.B = B; a.K = K; a.R = R; a.C = C; a.F0 = F0; a.warp = warp;
    EnvTab tab;
    if (int rc = check_common(a, envH, envW, tab)) return rc;
    IRB_REQUIRE(g_diff && g_spec, "render_sg_bwd: null upstream gradient");
    const bool pix = g_normal || g_rough;
    const bool lobe = g_axis || g_lamb || g_weight || g_albedo;
    if (!pix && !lobe) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int variant = tma_ok(a, true) ? render_variant(true, K) : 2;
    if (variant == 0 && tab.D % (2 * (K <= 6 ? 1 : (K <= 12 ? 2 : 4))) != 0) variant = 1;
    if (variant == 1) {
        const char* w = "render_sg_bwd(tma)";
        if (K <= 12) {
            if (pix && lobe) return launch_tma(render_sg_tma_kernel<1, 12, 12, true, true>, a, tab, true, st, w);
            if (pix) return launch_tma(render_sg_tma_kernel<1, 12, 12, true, false>, a, tab, true, st, w);
            return launch_tma(render_sg_tma_kernel<1, 12, 12, false, true>, a, tab, true, st, w);
        }
        if (pix && lobe) return launch_tma(render_sg_tma_kernel<1, 24, 12, true, true>, a, tab, true, st, w);
        if (pix) return launch_tma(render_sg_tma_kernel<1, 24, 12, true, false>, a, tab, true, st, w);
        return launch_tma(render_sg_tma_kernel<1, 24, 12, false, true>, a, tab, true, st, w);
    }
    if (variant == 0) {
        int rc;
        const char* w = "render_sg_bwd(group)";
        // saved: U_k / V_k of the forward pass are staged with the tile (16-byte alignment like every staged plane)
        const bool saved = lobe && a.lobe_uv_in != nullptr && (((uintptr_t)a.lobe_uv_in & 15u) == 0);
        if (!saved) a.lobe_uv_in = nullptr;
        // kernel = <1, TPP, PIX, LOBE, resident CTAs (2: 255 registers for the full backward, 3 otherwise), SAVED, FULLK>
#define IRB_BWD_K(TPP, PIXV, LOBEV, MINBV, SAVEDV)                                                                                   \
    (K == TPP * kLPT ? launch_group(render_sg_group_kernel<1, TPP, PIXV, LOBEV, MINBV, SAVEDV, true>, a, tab, true, st, w, kThreads / TPP, SAVEDV) \
                     : launch_group(render_sg_group_kernel<1, TPP, PIXV, LOBEV, MINBV, SAVEDV, false>, a, tab, true, st, w, kThreads / TPP, SAVEDV))
#define IRB_GROUP_BWD(TPP)                                                                                              \
    (pix && lobe ? (saved ? IRB_BWD_K(TPP, true, true, IRB_BWD_MINB, true) : IRB_BWD_K(TPP, true, true, IRB_BWD_MINB, false))                  \
     : pix       ? IRB_BWD_K(TPP, true, false, 3, false)                                                                 \
                 : (saved ? IRB_BWD_K(TPP, false, true, IRB_LOBEONLY_MINB, true) : IRB_BWD_K(TPP, false, true, IRB_LOBEONLY_MINB, false)))
        if (K <= 6) rc = IRB_GROUP_BWD(1);
        else if (K <= 12) rc = IRB_GROUP_BWD(2);
        else rc = IRB_GROUP_BWD(4);
#undef IRB_BWD_K
#undef IRB_GROUP_BWD
        return rc;   // pixels outside the regular regime were handled inside the kernel (group_general_pixel)
    }
    const dim3 grid((unsigned)(((size_t)R * C + kThreads - 1) / kThreads), (unsigned)B);
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
