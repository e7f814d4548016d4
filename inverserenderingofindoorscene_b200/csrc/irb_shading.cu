// irb_shading.cu -- SG lighting -> per-pixel diffuse irradiance ("shading") image, the GPU form of
// utils.predToShading (utils.py:156-195; caller testReal.py:636-643):
//     shading[c] = max( sum_d cos(El_d) sin(El_d) * sum_k W_kc exp(lam_k (a_k . w_d - 1)), 0 )
// over a DENSER hemisphere table than the render layer's (16 x 32 by default, global z-up frame, no BRDF).
// Inference-only (no backward in the reference).  One thread per pixel; lobes in chunks of 6 as packed f32x2 pairs
// (irb_pixel.cuh), the table is built once per CTA in shared memory from fp64 sin/cos (the reference builds it in
// fp64 NumPy).  MUFU-bound like the fused forward: 6 ex2 against 12 packed FMA-pipe instructions per direction.
#include "irb_common.cuh"
#include "irb_driver.cuh"

namespace irb {

constexpr int kShadeThreads = 128;
constexpr int kShadeMaxDirs = 2048;
constexpr int kShadeChunk = 6;     // lobes per register chunk (three f32x2 pairs)

__global__ void __launch_bounds__(kShadeThreads) sg_shading_kernel(const float* __restrict__ pred, float* __restrict__ out, int K,
                                                                   int R, int C, int envH, int envW) {
    extern __shared__ float4 tabs[];   // (w.x, w.y, w.z, cosEl sinEl)
    const int D = envH * envW;
    for (int i = threadIdx.x; i < D; i += kShadeThreads) {
        const int row = i / envW, col = i % envW;
        const double az = (((double)col + 0.5) / envW - 0.5) * 2.0 * 3.14159265358979323846;
        const double el = (((double)row + 0.5) / envH) * 3.14159265358979323846 / 2.0;
        double sa, ca, se, ce;
        sincos(az, &sa, &ca);
        sincos(el, &se, &ce);
        tabs[i] = make_float4((float)(se * ca), (float)(se * sa), (float)ce, (float)(ce * se));
    }
    __syncthreads();
    const size_t plane = (size_t)R * C;
    const size_t pix = (size_t)blockIdx.x * kShadeThreads + threadIdx.x;
    if (pix >= plane) return;
    const size_t b = blockIdx.y;
    // pred[b] = [axis (K,3) | lambOrig (K) | weightOrig (K,3)] x plane, exactly the layout of one image of the SG volume
    RenderSgArgs a{};
    const float* base = pred + b * (size_t)(7 * K) * plane;
    a.axis = base;
    a.lamb = base + (size_t)(3 * K) * plane;
    a.weight = base + (size_t)(4 * K) * plane;
    a.K = K;
    a.warp = 1;
    float sh[3] = {0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += kShadeChunk) {
        Lobes<float, kShadeChunk> L;
        load_lobes(L, a, 0, pix, plane, k0);
        LobesPk<kShadeChunk> Lp;
        pack_lobes<kShadeChunk>(Lp, L);
        pk2 U2[kShadeChunk / 2];
#pragma unroll
        for (int i = 0; i < kShadeChunk / 2; ++i) U2[i] = pk(0.f, 0.f);
        for (int d = 0; d < D; ++d) {
            const float4 t = tabs[d];
            pk2 arg[kShadeChunk / 2], e[kShadeChunk / 2];
            lobe_exps_pk<kShadeChunk>(Lp, pk(t.x, t.x), pk(t.y, t.y), pk(t.z, t.z), arg, e);
            const pk2 w2 = pk(t.w, t.w);
#pragma unroll
            for (int i = 0; i < kShadeChunk / 2; ++i) U2[i] = fma2(w2, e[i], U2[i]);
        }
#pragma unroll
        for (int i = 0; i < kShadeChunk / 2; ++i) {
            float u[2];
            upk(U2[i], u[0], u[1]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int kk = k0 + 2 * i + j;
                if (kk < K) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) sh[c] = fmaf(load_weight<float>(a, 0, pix, plane, kk, c), u[j], sh[c]);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(b * 3 + c) * plane + pix] = fmaxf(sh[c], 0.f);
}

}  // namespace irb

using namespace irb;

extern "C" int irb_sg_shading(const float* pred, float* out, int B, int K, int R, int C, int envH, int envW, void* stream) {
    IRB_REQUIRE(B > 0 && K > 0 && R > 0 && C > 0, "sg_shading: non-positive shape");
    IRB_REQUIRE(pred && out, "sg_shading: null pointer");
    IRB_REQUIRE(envH > 0 && envW > 0 && (long long)envH * envW <= kShadeMaxDirs, "sg_shading: need 0 < envH*envW <= %d",
                kShadeMaxDirs);
    IRB_REQUIRE(B <= 65535, "sg_shading: B too large");
    const size_t smem = sizeof(float4) * (size_t)envH * envW;
    const dim3 grid((unsigned)(((size_t)R * C + kShadeThreads - 1) / kShadeThreads), (unsigned)B);
    sg_shading_kernel<<<grid, kShadeThreads, smem, (cudaStream_t)stream>>>(pred, out, K, R, C, envH, envW);
    return check_launch("sg_shading");
}
