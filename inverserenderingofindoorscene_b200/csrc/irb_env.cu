// irb_env.cu -- UNFUSED boundary kernels over a materialised env-map [B,3,R,C,envH,envW]:
//   sg_to_env fwd/bwd   (models.output2env.fromSGtoIm / output2env, models.py:371-404)
//   render_env fwd/bwd  (models.renderingLayer.forwardEnv after pooling, models.py:466-520)
// Thread-per-pixel like the fused kernel; the env-map is read/written as 16-byte vectors along
// the direction axis (a pixel's 128 directions of one channel are one 512-byte run, so every
// 32-byte sector fetched is fully used by the thread that fetched it).  These are the exact
// drop-ins for callers that hand forwardEnv an arbitrary env tensor; the fused kernel in
// irb_render_sg.cu is the fast path when the env-map comes from SG lobes.
#include "irb_common.cuh"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "irb_driver_env.cuh"
#include "irb_tma.cuh"
#include "irb_launch.cuh"
#include "irb_sg_env_group.cuh"
#include "irb_render_env_tma.cuh"
#include "irb_sg_env_tile.cuh"

namespace irb {

constexpr int kThreadsEnv = 128;

template <int KC>
__global__ void __launch_bounds__(kThreadsEnv) sg_to_env_fwd_kernel(const SgEnvArgs s, const __grid_constant__ EnvTab tab) {
    const size_t pix = (size_t)blockIdx.x * kThreadsEnv + threadIdx.x;
    if (pix >= (size_t)s.R * s.C) return;
    drv_sg_to_env_fwd<float, KC>(s, tab, blockIdx.y, pix);
}
template <int KC>
__global__ void __launch_bounds__(kThreadsEnv) sg_to_env_bwd_kernel(const SgEnvArgs s, const __grid_constant__ EnvTab tab) {
    const size_t pix = (size_t)blockIdx.x * kThreadsEnv + threadIdx.x;
    if (pix >= (size_t)s.R * s.C) return;
    drv_sg_to_env_bwd<float, KC>(s, tab, blockIdx.y, pix);
}
__global__ void __launch_bounds__(kThreadsEnv) render_env_fwd_kernel(const RenderEnvArgs s, const __grid_constant__ EnvTab tab) {
    const size_t pix = (size_t)blockIdx.x * kThreadsEnv + threadIdx.x;
    if (pix >= (size_t)s.R * s.C) return;
    drv_render_env_fwd<float>(s, tab, blockIdx.y, pix);
}
__global__ void __launch_bounds__(kThreadsEnv) render_env_bwd_kernel(const RenderEnvArgs s, const __grid_constant__ EnvTab tab) {
    const size_t pix = (size_t)blockIdx.x * kThreadsEnv + threadIdx.x;
    if (pix >= (size_t)s.R * s.C) return;
    drv_render_env_bwd<float>(s, tab, blockIdx.y, pix);
}

// persistent TMA-staged lobe-split launch (irb_sg_env_group.cuh); falls back to the plain kernels when a plane is not
// 16-byte addressable, when K > 24, or with IRB_SGENV_VARIANT=plain
template <typename KernelT>
static int launch_sg_group(KernelT kernel, const SgEnvArgs& s, const EnvTab& tab, int tile, cudaStream_t st, const char* what) {
    const size_t plane = (size_t)s.R * s.C;
    const size_t tiles_per_img = (plane + tile - 1) / tile;
    const size_t total = tiles_per_img * (size_t)s.B;
    const size_t smem = 2 * (size_t)(7 * s.K) * tile * sizeof(float);
    int per_sm = 1;
    if (int e = persistent_ctas_per_sm(kernel, 128, smem, &per_sm)) {
        set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString((cudaError_t)e));
        return e;
    }
    size_t grid = (size_t)cached_sms() * per_sm;
    if (grid > total) grid = total;
    if (const char* cap = getenv("IRB_TMA_MAX_CTAS")) {
        const long c = atol(cap);
        if (c > 0 && (size_t)c < grid) grid = (size_t)c;
    }
    TileMaps maps;
    memset(&maps, 0, sizeof(maps));
    const char* tt = getenv("IRB_TMA_TENSOR");
    if (!(tt && tt[0] == '0') && plane < 0x7fffffffull) {
        const unsigned K = (unsigned)s.K, B = (unsigned)s.B, T = (unsigned)tile;
        maps.on = cached_tile_map(&maps.m[0], s.axis, plane, 3 * K, B, T) && cached_tile_map(&maps.m[1], s.lamb, plane, K, B, T) &&
                  cached_tile_map(&maps.m[2], s.weight, plane, 3 * K, B, T);
    }
    kernel<<<(unsigned)grid, 128, smem, st>>>(s, tab, maps, (unsigned)total, (unsigned)tiles_per_img);
    return check_launch(what);
}

static bool sg_group_ok(const SgEnvArgs& s, const void* extra) {
    const char* v = getenv("IRB_SGENV_VARIANT");
    if (v && v[0] == 'p') return false;
    if (s.K > 24) return false;
    const size_t plane = (size_t)s.R * s.C;
    if (plane % 4 != 0 || plane * (size_t)s.B >= 0xffffffffull) return false;
    const void* ptrs[] = {s.axis, s.lamb, s.weight, extra};
    for (const void* p : ptrs)
        if (((uintptr_t)p & 15u) != 0) return false;
    return true;
}

// IRB_DISABLE_TMA=1 / IRB_TMA_TENSOR=0: the tensor-map kernels of this file step aside too
static bool env_tiles_disabled() {
    const char* dis = getenv("IRB_DISABLE_TMA");
    const char* tt = getenv("IRB_TMA_TENSOR");
    return (dis && dis[0] == '1') || (tt && tt[0] == '0');
}

// TMA-staged render_env (irb_render_env_tma.cuh); returns -100 when the shape does not qualify (caller takes the plain kernel)
template <bool BWD>
static int launch_render_env_tma(const RenderEnvArgs& s, const EnvTab& tab, cudaStream_t st, const char* what) {
    const char* v = getenv("IRB_RENDER_ENV_VARIANT");
    if (v && v[0] == 'p') return -100;
    if (env_tiles_disabled()) return -100;
    const size_t plane = (size_t)s.R * s.C;
    constexpr unsigned DC = ReCfg<BWD>::DC;
    if (tab.D % DC != 0 || plane >= 0x7fffffffull) return -100;
    EnvTileMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (!cached_env_tile_map(&maps.env, s.env, plane, (unsigned)tab.D, (size_t)s.B * 3, DC)) return -100;
    if (BWD && s.g_env && !cached_env_tile_map(&maps.genv, s.g_env, plane, (unsigned)tab.D, (size_t)s.B * 3, DC)) return -100;
    const size_t tiles_per_img = (plane + RE_PT - 1) / RE_PT;
    const size_t total = tiles_per_img * (size_t)s.B;
    if (total >= 0x7fffffffull) return -100;
    const size_t smem = (size_t)RE_S * ReCfg<BWD>::STAGE_BYTES + RE_S * (sizeof(uint64_t) + sizeof(unsigned));
    int per_sm = 1;
    if (int e = persistent_ctas_per_sm(render_env_tma_kernel<BWD>, RE_PT, smem, &per_sm)) {
        set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString((cudaError_t)e));
        return e;
    }
    size_t grid = (size_t)cached_sms() * per_sm;
    if (grid > total) grid = total;
    if (const char* cap = getenv("IRB_TMA_MAX_CTAS")) {
        const long c = atol(cap);
        if (c > 0 && (size_t)c < grid) grid = (size_t)c;
    }
    render_env_tma_kernel<BWD><<<(unsigned)grid, RE_PT, smem, st>>>(s, tab, maps, (unsigned)total, (unsigned)tiles_per_img);
    return check_launch(what);
}

// thread-per-pixel, TMA-stored sg_to_env forward (irb_sg_env_tile.cuh) for K = 6 / 12; -100 = does not qualify
template <int KC>
static int launch_sg_env_tile_fwd(const SgEnvArgs& s, const EnvTab& tab, cudaStream_t st, const char* what) {
    const size_t plane = (size_t)s.R * s.C;
    constexpr unsigned DC = IRB_SGT_DC;
    if (tab.D % DC != 0 || plane >= 0x7fffffffull || env_tiles_disabled()) return -100;
    CUtensorMap emap;
    memset(&emap, 0, sizeof(emap));
    if (!cached_env_tile_map(&emap, s.env, plane, (unsigned)tab.D, (size_t)s.B * 3, DC)) return -100;
    const size_t tiles_per_img = (plane + RE_PT - 1) / RE_PT;
    const size_t total = tiles_per_img * (size_t)s.B;
    if (total >= 0x7fffffffull) return -100;
    const size_t smem = (size_t)IRB_SGT_STAGES * 3 * RE_PT * DC * 4 + IRB_SGT_STAGES * (sizeof(uint64_t) + sizeof(unsigned));
    int per_sm = 1;
    if (int e = persistent_ctas_per_sm(sg_env_tile_fwd_kernel<KC>, RE_PT, smem, &per_sm)) {
        set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString((cudaError_t)e));
        return e;
    }
    size_t grid = (size_t)cached_sms() * per_sm;
    if (grid > total) grid = total;
    if (const char* cap = getenv("IRB_TMA_MAX_CTAS")) {
        const long c = atol(cap);
        if (c > 0 && (size_t)c < grid) grid = (size_t)c;
    }
    sg_env_tile_fwd_kernel<KC><<<(unsigned)grid, RE_PT, smem, st>>>(s, tab, emap, (unsigned)total, (unsigned)tiles_per_img);
    return check_launch(what);
}

// thread-per-pixel reconstruction statistics (MODE 2) / loss (MODE 3) with the ground truth staged by TMA; -100 = not taken
template <int KC, int MODE>
static int launch_sg_env_tile_stat(const SgEnvArgs& s, const EnvTab& tab, cudaStream_t st, const char* what) {
    const size_t plane = (size_t)s.R * s.C;
    constexpr unsigned DC = 8, S = 3;
    if (tab.D % DC != 0 || plane >= 0x7fffffffull || env_tiles_disabled() || ((uintptr_t)s.env_gt & 15u)) return -100;
    CUtensorMap gmap;
    memset(&gmap, 0, sizeof(gmap));
    if (!cached_env_tile_map(&gmap, s.env_gt, plane, (unsigned)tab.D, (size_t)s.B * 3, DC)) return -100;
    const size_t tiles_per_img = (plane + RE_PT - 1) / RE_PT;
    const size_t total = tiles_per_img * (size_t)s.B;
    if (total >= 0x7fffffffull) return -100;
    const size_t smem = (size_t)S * 3 * RE_PT * DC * 4 + S * (sizeof(uint64_t) + sizeof(unsigned));
    int per_sm = 1;
    if (int e = persistent_ctas_per_sm(sg_env_tile_stat_kernel<KC, MODE>, RE_PT, smem, &per_sm)) {
        set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString((cudaError_t)e));
        return e;
    }
    size_t grid = (size_t)cached_sms() * per_sm;
    if (grid > total) grid = total;
    if (const char* cap = getenv("IRB_TMA_MAX_CTAS")) {
        const long c = atol(cap);
        if (c > 0 && (size_t)c < grid) grid = (size_t)c;
    }
    sg_env_tile_stat_kernel<KC, MODE><<<(unsigned)grid, RE_PT, smem, st>>>(s, tab, gmap, (unsigned)total, (unsigned)tiles_per_img);
    return check_launch(what);
}
template <int MODE>
static int try_sg_env_tile_stat(const SgEnvArgs& s, const EnvTab& tab, cudaStream_t st, const char* what) {
    const char* v = getenv("IRB_SGENV_VARIANT");
    if (v && (v[0] == 'p' || v[0] == 'g')) return -100;
    if (s.K == 12) return launch_sg_env_tile_stat<12, MODE>(s, tab, st, what);
    if (s.K == 6) return launch_sg_env_tile_stat<6, MODE>(s, tab, st, what);
    return -100;
}

static int mk_tab(EnvTab& tab, int envH, int envW, int B, const char* who) {
    if (!make_env_tab(tab, envH, envW)) {
        set_error("%s: env table %dx%d unsupported (need envH*envW <= %d, multiple of 4)", who, envH, envW, IRB_MAX_DIRS);
        return IRB_E_UNSUPPORTED;
    }
    if (B > 65535) {
        set_error("%s: B=%d exceeds grid.y", who, B);
        return IRB_E_UNSUPPORTED;
    }
    return 0;
}

}  // namespace irb

using namespace irb;

extern "C" int irb_sg_to_env_fwd(const float* axis, const float* lamb, const float* weight, float* env, float* lamb_w,
                                 float* weight_w, int B, int K, int R, int C, int envH, int envW, int warp, void* stream) {
    IRB_REQUIRE(B > 0 && K > 0 && R > 0 && C > 0, "sg_to_env_fwd: non-positive shape");
    IRB_REQUIRE(axis && lamb && weight && env, "sg_to_env_fwd: null pointer");
    IRB_REQUIRE(((uintptr_t)env & 15u) == 0, "sg_to_env_fwd: env must be 16-byte aligned (it is written as 16-byte vectors)");
    EnvTab tab;
    if (int rc = mk_tab(tab, envH, envW, B, "sg_to_env_fwd")) return rc;
    SgEnvArgs s{};
    s.axis = axis; s.lamb = lamb; s.weight = weight; s.env = env; s.lamb_w = lamb_w; s.weight_w = weight_w;
    s.B = B; s.K = K; s.R = R; s.C = C; s.warp = warp;
    {
        const char* v = getenv("IRB_SGENV_VARIANT");
        if (!(v && (v[0] == 'p' || v[0] == 'g')) && (K == 6 || K == 12)) {
            const int rc = K == 12 ? launch_sg_env_tile_fwd<12>(s, tab, (cudaStream_t)stream, "sg_to_env_fwd(tile)")
                                   : launch_sg_env_tile_fwd<6>(s, tab, (cudaStream_t)stream, "sg_to_env_fwd(tile)");
            if (rc != -100) return rc;
        }
    }
    if (sg_group_ok(s, env)) {
        cudaStream_t st = (cudaStream_t)stream;
        const char* w = "sg_to_env_fwd(group)";
        if (K <= 6) return launch_sg_group(sg_env_group_kernel<0, 1>, s, tab, 128, st, w);
        if (K <= 12) return launch_sg_group(sg_env_group_kernel<0, 2>, s, tab, 64, st, w);
        return launch_sg_group(sg_env_group_kernel<0, 4>, s, tab, 32, st, w);
    }
    const dim3 grid((unsigned)(((size_t)R * C + kThreadsEnv - 1) / kThreadsEnv), (unsigned)B);
    sg_to_env_fwd_kernel<12><<<grid, kThreadsEnv, 0, (cudaStream_t)stream>>>(s, tab);
    return check_launch("sg_to_env_fwd");
}

extern "C" int irb_sg_to_env_bwd(const float* axis, const float* lamb, const float* weight, const float* g_env,
                                 float* g_axis, float* g_lamb, float* g_weight, int B, int K, int R, int C, int envH,
                                 int envW, int warp, void* stream) {
    IRB_REQUIRE(B > 0 && K > 0 && R > 0 && C > 0, "sg_to_env_bwd: non-positive shape");
    IRB_REQUIRE(axis && lamb && weight && g_env, "sg_to_env_bwd: null pointer");
    IRB_REQUIRE(((uintptr_t)g_env & 15u) == 0, "sg_to_env_bwd: g_env must be 16-byte aligned (it is read as 16-byte vectors)");
    EnvTab tab;
    if (int rc = mk_tab(tab, envH, envW, B, "sg_to_env_bwd")) return rc;
    SgEnvArgs s{};
    s.axis = axis; s.lamb = lamb; s.weight = weight; s.g_env = g_env;
    s.g_axis = g_axis; s.g_lamb = g_lamb; s.g_weight = g_weight;
    s.B = B; s.K = K; s.R = R; s.C = C; s.warp = warp;
    if (sg_group_ok(s, g_env)) {
        cudaStream_t st = (cudaStream_t)stream;
        const char* w = "sg_to_env_bwd(group)";
        if (K <= 6) return launch_sg_group(sg_env_group_kernel<1, 1>, s, tab, 128, st, w);
        if (K <= 12) return launch_sg_group(sg_env_group_kernel<1, 2>, s, tab, 64, st, w);
        return launch_sg_group(sg_env_group_kernel<1, 4>, s, tab, 32, st, w);
    }
    const dim3 grid((unsigned)(((size_t)R * C + kThreadsEnv - 1) / kThreadsEnv), (unsigned)B);
    sg_to_env_bwd_kernel<12><<<grid, kThreadsEnv, 0, (cudaStream_t)stream>>>(s, tab);
    return check_launch("sg_to_env_bwd");
}

// Fused reconstruction loss of the light branch (wrapperBRDFLight.py:177-188): the predicted env-map is evaluated from
// its SG lobes inside the kernels and compared with the ground truth on the fly, so the only large stream is the
// ground-truth env-map (read once per pass).  stats[b] = {sum E G s^2, sum E^2 s^2, sum dlog^2 s, sum s} in fp64.
extern "C" int irb_sg_reconst_fwd(const float* axis, const float* lamb, const float* weight, const float* env_gt,
                                  const float* seg, double* stats, float* g_axis, float* g_lamb, float* g_weight, int B, int K,
                                  int R, int C, int envH, int envW, int warp, float offset, void* stream) {
    IRB_REQUIRE(B > 0 && K > 0 && R > 0 && C > 0, "sg_reconst_fwd: non-positive shape");
    IRB_REQUIRE(axis && lamb && weight && env_gt && seg && stats, "sg_reconst_fwd: null pointer");
    IRB_REQUIRE(offset > 0.f, "sg_reconst_fwd: offset must be positive");
    EnvTab tab;
    if (int rc = mk_tab(tab, envH, envW, B, "sg_reconst_fwd")) return rc;
    SgEnvArgs s{};
    s.axis = axis; s.lamb = lamb; s.weight = weight; s.env_gt = env_gt; s.seg = seg; s.stats = stats; s.offset = offset;
    s.g_axis = g_axis; s.g_lamb = g_lamb; s.g_weight = g_weight;
    s.B = B; s.K = K; s.R = R; s.C = C; s.warp = warp;
    const bool with_grads = g_axis || g_lamb || g_weight;   // second pass = loss + unit-gradient in one sweep
    if (!sg_group_ok(s, env_gt)) {
        set_error("sg_reconst_fwd: needs K <= 24, R*C %% 4 == 0 and 16-byte aligned planes");
        return IRB_E_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 4 * (size_t)B, st);
    if (e != cudaSuccess) {
        set_error("sg_reconst_fwd: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    int rc = try_sg_env_tile_stat<2>(s, tab, st, "sg_reconst_fwd(stats,tile)");
    if (rc != -100) {   // statistics done by the tile kernel; second pass
        if (rc) return rc;
        if (!with_grads) {
            rc = try_sg_env_tile_stat<3>(s, tab, st, "sg_reconst_fwd(loss,tile)");
            if (rc != -100) return rc;
        }
        if (K <= 6) return with_grads ? launch_sg_group(sg_env_group_kernel<5, 1>, s, tab, 128, st, "sg_reconst_fwd(loss+grad)")
                                      : launch_sg_group(sg_env_group_kernel<3, 1>, s, tab, 128, st, "sg_reconst_fwd(loss)");
        return with_grads ? launch_sg_group(sg_env_group_kernel<5, 2>, s, tab, 64, st, "sg_reconst_fwd(loss+grad)")
                          : launch_sg_group(sg_env_group_kernel<3, 2>, s, tab, 64, st, "sg_reconst_fwd(loss)");
    }
    if (K <= 6) {
        rc = launch_sg_group(sg_env_group_kernel<2, 1>, s, tab, 128, st, "sg_reconst_fwd(stats)");
        if (!rc) rc = with_grads ? launch_sg_group(sg_env_group_kernel<5, 1>, s, tab, 128, st, "sg_reconst_fwd(loss+grad)")
                                 : launch_sg_group(sg_env_group_kernel<3, 1>, s, tab, 128, st, "sg_reconst_fwd(loss)");
    } else if (K <= 12) {
        rc = launch_sg_group(sg_env_group_kernel<2, 2>, s, tab, 64, st, "sg_reconst_fwd(stats)");
        if (!rc) rc = with_grads ? launch_sg_group(sg_env_group_kernel<5, 2>, s, tab, 64, st, "sg_reconst_fwd(loss+grad)")
                                 : launch_sg_group(sg_env_group_kernel<3, 2>, s, tab, 64, st, "sg_reconst_fwd(loss)");
    } else {
        rc = launch_sg_group(sg_env_group_kernel<2, 4>, s, tab, 32, st, "sg_reconst_fwd(stats)");
        if (!rc) rc = with_grads ? launch_sg_group(sg_env_group_kernel<5, 4>, s, tab, 32, st, "sg_reconst_fwd(loss+grad)")
                                 : launch_sg_group(sg_env_group_kernel<3, 4>, s, tab, 32, st, "sg_reconst_fwd(loss)");
    }
    return rc;
}

extern "C" int irb_sg_reconst_bwd(const float* axis, const float* lamb, const float* weight, const float* env_gt,
                                  const float* seg, const double* stats, const float* g_num, float* g_axis, float* g_lamb,
                                  float* g_weight, int B, int K, int R, int C, int envH, int envW, int warp, float offset,
                                  void* stream) {
    IRB_REQUIRE(B > 0 && K > 0 && R > 0 && C > 0, "sg_reconst_bwd: non-positive shape");
    IRB_REQUIRE(axis && lamb && weight && env_gt && seg && stats && g_num, "sg_reconst_bwd: null pointer");
    IRB_REQUIRE(offset > 0.f, "sg_reconst_bwd: offset must be positive");
    EnvTab tab;
    if (int rc = mk_tab(tab, envH, envW, B, "sg_reconst_bwd")) return rc;
    SgEnvArgs s{};
    s.axis = axis; s.lamb = lamb; s.weight = weight; s.env_gt = env_gt; s.seg = seg;
    s.stats = const_cast<double*>(stats); s.g_num = g_num; s.offset = offset;
    s.g_axis = g_axis; s.g_lamb = g_lamb; s.g_weight = g_weight;
    s.B = B; s.K = K; s.R = R; s.C = C; s.warp = warp;
    if (!sg_group_ok(s, env_gt)) {
        set_error("sg_reconst_bwd: needs K <= 24, R*C %% 4 == 0 and 16-byte aligned planes");
        return IRB_E_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 6) return launch_sg_group(sg_env_group_kernel<4, 1>, s, tab, 128, st, "sg_reconst_bwd");
    if (K <= 12) return launch_sg_group(sg_env_group_kernel<4, 2>, s, tab, 64, st, "sg_reconst_bwd");
    return launch_sg_group(sg_env_group_kernel<4, 4>, s, tab, 32, st, "sg_reconst_bwd");
}

extern "C" int irb_render_env_fwd(const float* albedo, const float* normal, const float* rough, const float* env,
                                  const float* view, float* out_diff, float* out_spec, int B, int R, int C, int envH,
                                  int envW, float F0, void* stream) {
    IRB_REQUIRE(B > 0 && R > 0 && C > 0, "render_env_fwd: non-positive shape");
    IRB_REQUIRE(albedo && normal && rough && env && view && out_diff && out_spec, "render_env_fwd: null pointer");
    IRB_REQUIRE(((uintptr_t)env & 15u) == 0, "render_env_fwd: env must be 16-byte aligned (it is read as 16-byte vectors)");
    EnvTab tab;
    if (int rc = mk_tab(tab, envH, envW, B, "render_env_fwd")) return rc;
    RenderEnvArgs s{};
    s.albedo = albedo; s.normal = normal; s.rough = rough; s.env = env; s.view = view;
    s.out_diff = out_diff; s.out_spec = out_spec;
    s.B = B; s.R = R; s.C = C; s.F0 = F0;
    {
        const int rc = launch_render_env_tma<false>(s, tab, (cudaStream_t)stream, "render_env_fwd(tma)");
        if (rc != -100) return rc;
    }
    const dim3 grid((unsigned)(((size_t)R * C + kThreadsEnv - 1) / kThreadsEnv), (unsigned)B);
    render_env_fwd_kernel<<<grid, kThreadsEnv, 0, (cudaStream_t)stream>>>(s, tab);
    return check_launch("render_env_fwd");
}

extern "C" int irb_render_env_bwd(const float* albedo, const float* normal, const float* rough, const float* env,
                                  const float* view, const float* g_diff, const float* g_spec, float* g_albedo,
                                  float* g_normal, float* g_rough, float* g_env, int B, int R, int C, int envH, int envW,
                                  float F0, void* stream) {
    IRB_REQUIRE(B > 0 && R > 0 && C > 0, "render_env_bwd: non-positive shape");
    IRB_REQUIRE(albedo && normal && rough && env && view && g_diff && g_spec, "render_env_bwd: null pointer");
    IRB_REQUIRE((((uintptr_t)env | (uintptr_t)g_env) & 15u) == 0, "render_env_bwd: env / g_env must be 16-byte aligned");
    EnvTab tab;
    if (int rc = mk_tab(tab, envH, envW, B, "render_env_bwd")) return rc;
    RenderEnvArgs s{};
    s.albedo = albedo; s.normal = normal; s.rough = rough; s.env = env; s.view = view;
    s.g_diff = g_diff; s.g_spec = g_spec;
    s.g_albedo = g_albedo; s.g_normal = g_normal; s.g_rough = g_rough; s.g_env = g_env;
    s.B = B; s.R = R; s.C = C; s.F0 = F0;
    {
        const int rc = launch_render_env_tma<true>(s, tab, (cudaStream_t)stream, "render_env_bwd(tma)");
        if (rc != -100) return rc;
    }
    const dim3 grid((unsigned)(((size_t)R * C + kThreadsEnv - 1) / kThreadsEnv), (unsigned)B);
    render_env_bwd_kernel<<<grid, kThreadsEnv, 0, (cudaStream_t)stream>>>(s, tab);
    return check_launch("render_env_bwd");
}
