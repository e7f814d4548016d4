// irb_pool.cu -- F.adaptive_avg_pool2d forward/backward as used by renderingLayer.forwardEnv
// (models.py:465-469): window of output row i is [floor(i*H/R), ceil((i+1)*H/R)), likewise columns.
// One thread per OUTPUT element (fwd) / per INPUT element (bwd, gathers the windows covering it,
// so no atomics).  Memory-bound elementwise work; consecutive threads walk consecutive columns.
#include "irb_common.cuh"

namespace irb {

__device__ __forceinline__ int win_lo(int i, int in, int out) { return (int)(((long long)i * in) / out); }
__device__ __forceinline__ int win_hi(int i, int in, int out) { return (int)((((long long)(i + 1)) * in + out - 1) / out); }

__global__ void avgpool_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int R, int C) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)N * R * C;
    if (t >= total) return;
    const int x = (int)(t % C), r = (int)((t / C) % R);
    const size_t n = t / ((size_t)R * C);
    const int h0 = win_lo(r, H, R), h1 = win_hi(r, H, R), w0 = win_lo(x, W, C), w1 = win_hi(x, W, C);
    const float* p = in + n * (size_t)H * W;
    float acc = 0.f;
    for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) acc += __ldg(p + (size_t)h * W + w);
    out[t] = acc / (float)((h1 - h0) * (w1 - w0));
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ g_out, float* __restrict__ g_in, int N, int H, int W, int R, int C) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)N * H * W;
    if (t >= total) return;
    const int w = (int)(t % W), h = (int)((t / W) % H);
    const size_t n = t / ((size_t)H * W);
    // candidate output rows/cols whose window may contain (h, w); each is verified below
    int r_lo = (int)(((long long)h * R) / H) - 1;
    if (r_lo < 0) r_lo = 0;
    int r_hi = (int)((((long long)h + 1) * R + H - 1) / H);
    if (r_hi > R - 1) r_hi = R - 1;
    int c_lo = (int)(((long long)w * C) / W) - 1;
    if (c_lo < 0) c_lo = 0;
    int c_hi = (int)((((long long)w + 1) * C + W - 1) / W);
    if (c_hi > C - 1) c_hi = C - 1;
    const float* g = g_out + n * (size_t)R * C;
    float acc = 0.f;
    for (int r = r_lo; r <= r_hi; ++r) {
        const int h0 = win_lo(r, H, R), h1 = win_hi(r, H, R);
        if (h < h0 || h >= h1) continue;
        for (int x = c_lo; x <= c_hi; ++x) {
            const int w0 = win_lo(x, W, C), w1 = win_hi(x, W, C);
            if (w < w0 || w >= w1) continue;
            acc += __ldg(g + (size_t)r * C + x) / (float)((h1 - h0) * (w1 - w0));
        }
    }
    g_in[t] = acc;
}

}  // namespace irb

using namespace irb;

extern "C" int irb_avgpool_fwd(const float* in, float* out, int N, int H, int W, int R, int C, void* stream) {
    IRB_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && R > 0 && C > 0, "avgpool_fwd: bad argument");
    const size_t total = (size_t)N * R * C;
    avgpool_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, N, H, W, R, C);
    return check_launch("avgpool_fwd");
}

extern "C" int irb_avgpool_bwd(const float* g_out, float* g_in, int N, int H, int W, int R, int C, void* stream) {
    IRB_REQUIRE(g_out && g_in && N > 0 && H > 0 && W > 0 && R > 0 && C > 0, "avgpool_bwd: bad argument");
    const size_t total = (size_t)N * H * W;
    avgpool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_out, g_in, N, H, W, R, C);
    return check_launch("avgpool_bwd");
}
