// pipe_rates.cu -- measures the FP32 issue rates that bound the fused render kernel on B200:
// FFMA (register forms), FFMA2 (fma.rn.f32x2), MUFU.EX2, and the fwd/bwd instruction mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates.bin pipe_rates.cu
// Output: one line per test: ops per clock per SM (SM clock from clock64 deltas).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long d, a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

template <int MODE> __global__ void k(float* out, long long* cyc, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    float p = a, q = b;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // 16 independent FFMA, x = x*p + q  (2 distinct + dest)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], p, q);
        } else if (MODE == 1) {  // accumulate form acc += p * x[j] : 3 distinct regs
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(p, x[(i + 1) & 15], x[i]);
        } else if (MODE == 2) {  // FFMA2, 8 pairs
#pragma unroll
            for (int i = 0; i < 16; i += 2) ffma2(x[i], x[i + 1], p, q, x[i], x[i + 1]);
        } else if (MODE == 3) {  // MUFU.EX2 only
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
        } else if (MODE == 4) {  // fwd mix: per lobe 5 FFMA-pipe + 1 MUFU
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float d = fmaf(x[i], p, fmaf(x[i], q, fmaf(x[i], a, -1.0f)));
                float e = ex2(d * q);
                x[i] = fmaf(p, e, x[i]);
                x[i + 1] = fmaf(q, e, x[i + 1]);
            }
        } else if (MODE == 5) {  // immediate-form FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], 1.0001f, 0.5f);
        } else if (MODE == 6) {  // FMUL + FADD alternating
#pragma unroll
            for (int i = 0; i < 16; i += 2) { x[i] = x[i] * p; x[i + 1] = x[i + 1] + q; }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, double ops_per_iter_thread, int warps_per_sm) {
    int nsm = 148;
    int threads = 256, blocks = nsm * (warps_per_sm * 32 / threads);
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.9999f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.9999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    double ops_sm = ops_per_iter_thread * ITERS * (double)(warps_per_sm * 32);
    printf("%-34s warps/SM=%2d  %8.1f ops/clk/SM   (%.0f cyc, %.3f ms, eff clk %.0f MHz)\n", name, warps_per_sm,
           ops_sm / avg, avg, ms, avg / (ms * 1e3));
    cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
    for (int w : {8, 16, 32}) {
        run<0>("FFMA x=x*p+q (2 src regs)", 16, w);
        run<1>("FFMA acc+=p*x (3 src regs)", 16, w);
        run<5>("FFMA imm form", 16, w);
        run<2>("FFMA2 (counted as 2 FMA each)", 16, w);
        run<6>("FMUL/FADD", 16, w);
        run<3>("MUFU.EX2", 16, w);
        run<4>("fwd mix (5 FMA-pipe+1 MUFU)/lobe: FMA-ops", 8 * 6, w);
    }
    return 0;
}
