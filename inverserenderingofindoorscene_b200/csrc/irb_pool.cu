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

// the three BRDF maps of forwardEnv (albedo 3, normal 3, rough 1 channels) in ONE launch: blockIdx.y selects the tensor
struct PoolSet {
    const float* in[3];
    float* out[3];
    int n[3];   // images * channels of each tensor
};
__global__ void avgpool3_fwd_kernel(const PoolSet ps, int H, int W, int R, int C) {
    const int k = blockIdx.y;
    const float* in = ps.in[k];
    float* out = ps.out[k];
    if (!in || !out) return;
    const size_t total = (size_t)ps.n[k] * R * C;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % C), r = (int)((t / C) % R);
        const size_t n = t / ((size_t)R * C);
        const int h0 = win_lo(r, H, R), h1 = win_hi(r, H, R), w0 = win_lo(x, W, C), w1 = win_hi(x, W, C);
        const float* p = in + n * (size_t)H * W;
        float acc = 0.f;
        for (int h = h0; h < h1; ++h)
            for (int w = w0; w < w1; ++w) acc += __ldg(p + (size_t)h * W + w);
        out[t] = acc / (float)((h1 - h0) * (w1 - w0));
    }
}
__global__ void avgpool3_bwd_kernel(const PoolSet ps, int H, int W, int R, int C) {
    const int k = blockIdx.y;
    const float* g_out = ps.in[k];
    float* g_in = ps.out[k];
    if (!g_out || !g_in) return;
    const size_t total = (size_t)ps.n[k] * H * W;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(t % W), h = (int)((t / W) % H);
        const size_t n = t / ((size_t)H * W);
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
}

}  // namespace irb

using namespace irb;

// forwardEnv pools albedo / normal / rough together (models.py:465-469): one launch instead of three; any pair may be NULL
extern "C" int irb_avgpool3_fwd(const float* in0, const float* in1, const float* in2, float* out0, float* out1, float* out2,
                                int n0, int n1, int n2, int H, int W, int R, int C, void* stream) {
    IRB_REQUIRE(H > 0 && W > 0 && R > 0 && C > 0 && n0 >= 0 && n1 >= 0 && n2 >= 0, "avgpool3_fwd: bad argument");
    PoolSet ps{{in0, in1, in2}, {out0, out1, out2}, {n0, n1, n2}};
    const int nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    if (nmax == 0) return 0;
    size_t gx = ((size_t)nmax * R * C + 255) / 256;
    if (gx > 148 * 16) gx = 148 * 16;
    avgpool3_fwd_kernel<<<dim3((unsigned)gx, 3), 256, 0, (cudaStream_t)stream>>>(ps, H, W, R, C);
    return check_launch("avgpool3_fwd");
}
// in = gradients at R x C, out = gradients at H x W
extern "C" int irb_avgpool3_bwd(const float* g0, const float* g1, const float* g2, float* gi0, float* gi1, float* gi2, int n0,
                                int n1, int n2, int H, int W, int R, int C, void* stream) {
    IRB_REQUIRE(H > 0 && W > 0 && R > 0 && C > 0 && n0 >= 0 && n1 >= 0 && n2 >= 0, "avgpool3_bwd: bad argument");
    PoolSet ps{{g0, g1, g2}, {gi0, gi1, gi2}, {n0, n1, n2}};
    const int nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    if (nmax == 0) return 0;
    size_t gx = ((size_t)nmax * H * W + 255) / 256;
    if (gx > 148 * 16) gx = 148 * 16;
    avgpool3_bwd_kernel<<<dim3((unsigned)gx, 3), 256, 0, (cudaStream_t)stream>>>(ps, H, W, R, C);
    return check_launch("avgpool3_bwd");
}

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
