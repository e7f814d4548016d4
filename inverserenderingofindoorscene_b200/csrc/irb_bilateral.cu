// irb_bilateral.cu -- bilateral-grid solver on the GPU: grid build (hash, sort, unique, neighbour
// table), bistochastisation, Jacobi-preconditioned CG solve and its adjoint, splat/slice.
// Replaces BilateralGrid.py:44-207 and the per-image NumPy bridge BilateralLayer.py:24-124.
//
// Data layout (all in the caller-owned workspace, per image strided by P = H*W):
//   idx[P]      vertex of each pixel = rank of its hash among the image's sorted unique hashes
//   sp[P]       pixels sorted by (hash, pixel)  -> splat is a segmented, ORDERED sum (no atomics;
//               same summation order as the reference's CSR S.dot(x), hence bit-identical in fp64)
//   vstart[V+1] segment starts, ukeys[V] sorted unique hashes, nbr[10][P] (+-1 neighbour per dim, -1 = none)
//   n[P], m[P]  bistochastisation vectors (fp64);  solver vectors x,r,p,q,b (fp64, per channel)
// Arithmetic is fp64 like the reference (NumPy/SciPy doubles); the vertex indexing is integer-exact.
// Per-iteration structure of PCG: 3 kernels (p-update writing n*p channel-interleaved; mat-vec + <p,q>;
// x/r update + <r,z>, <r,r>),
// all images and channels batched in each launch, scalars accumulated per (image, channel, iteration).
#include <stdint.h>
#include <stdlib.h>

#include "irb_common.cuh"

namespace irb {

// resident 256-thread CTAs per SM the three PCG kernels are compiled for.  They are latency-bound streaming / gather kernels
// (ncu r02: 55-85 % long-scoreboard stalls), so occupancy is memory-level parallelism: measured on one B200, config 4
// fwd+bwd: unconstrained (44-58 registers, 4-5 CTAs) 8.06 ms, 6 CTAs (40 registers, 8-byte spills in the 3-channel
// kernels) 6.72 ms, 8 CTAs (32 registers, real spills) 6.79 ms.
#ifndef IRB_BG_MINB
#define IRB_BG_MINB 6
#endif
constexpr int BG_DIM = 5;
constexpr int BG_MAX_ITERS = 256;
constexpr int BG_MAX_CH = 4;

struct BgLayout {
    int B, H, W, ch;
    size_t P;
    size_t hdr, oor, idx, vstart, ukeys, nbr, n0, n1, m, m0, keysA, keysB, payA, payB, hist, dtot;
    size_t wsplat, minv, bvec, r, p0, p1, q, sol, np, scal, total;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static BgLayout make_layout(int B, int H, int W, int ch) {
    BgLayout L{};
    L.B = B; L.H = H; L.W = W; L.ch = ch;
    L.P = (size_t)H * W;
    const size_t P = L.P, Bz = (size_t)B;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t r = o; o = align_up(o + bytes); return r; };
    L.hdr = take(sizeof(int) * Bz);
    L.oor = take(sizeof(int) * Bz);   // per image: 1 if a grid coordinate left [0, 255) (hash aliasing; see irb_bg_range_flags)
    L.idx = take(4 * P * Bz);
    L.vstart = take(4 * (P + 1) * Bz);
    L.ukeys = take(8 * P * Bz);
    L.nbr = take(4 * 2 * BG_DIM * P * Bz);
    L.n0 = take(8 * P * Bz);
    L.n1 = take(8 * P * Bz);
    L.m = take(8 * P * Bz);
    L.m0 = take(8 * P * Bz);
    L.keysA = take(8 * P * Bz);
    L.keysB = take(8 * P * Bz);
    L.payA = take(4 * P * Bz);
    L.payB = take(4 * P * Bz);
    L.hist = take(4 * 256 * (size_t)256 * Bz);  // radix-sort bucket counters [img][digit][warp-chunk]
    L.dtot = take(4 * 256 * (size_t)8 * Bz);    // per-digit totals of every pass [pass][img][digit]
    L.wsplat = take(8 * P * Bz);
    L.minv = take(8 * P * Bz);
    const size_t vc = 8 * P * Bz * (size_t)(ch > 0 ? ch : 1);
    L.bvec = take(vc);
    L.r = take(vc);
    L.p0 = take(vc);
    L.p1 = take(vc);
    L.q = take(vc);
    L.sol = take(vc);  // adjoint solution (backward)
    L.np = take(8 * 4 * P * Bz);  // n_v * p_v, interleaved [vertex][4 channels]: one 32-byte sector per neighbour gather
    L.scal = take(8 * Bz * BG_MAX_CH * (BG_MAX_ITERS + 2) * 4);
    L.total = o;
    return L;
}

template <typename T> __host__ __device__ inline T* at(void* ws, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off); }
template <typename T> __host__ __device__ inline const T* at(const void* ws, size_t off) {
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(ws) + off);
}

// ---------------------------------------------------------------------------------------------
// 1. per-pixel hash (BilateralGrid.py:23-24, 44-59, 86-88)
// ---------------------------------------------------------------------------------------------
__global__ void bg_keys_kernel(const float* __restrict__ guide, float gscale, void* ws, BgLayout L, double sl, double sc, double ss) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.P * L.B) return;
    const size_t b = t / L.P, p = t % L.P;
    const int y = (int)(p / L.W), x = (int)(p % L.W);
    const float* g = guide + b * 3 * L.P + p;
    // BilateralLayer.py:52: curr_image * 255.0 is a float32 product
    const double r = (double)__fmul_rn(g[0], gscale), gg = (double)__fmul_rn(g[L.P], gscale),
                 bl = (double)__fmul_rn(g[2 * L.P], gscale);
    // np.tensordot -> OpenBLAS dgemm: fma(b, m2, fma(g, m1, r*m0)), then + offset
    const double Y = fma(bl, 0.114, fma(gg, 0.587, r * 0.299)) + 0.0;
    const double U = fma(bl, 0.5, fma(gg, -0.331264, r * -0.168736)) + 128.0;
    const double V = fma(bl, -0.081312, fma(gg, -0.418688, r * 0.5)) + 128.0;
    // (coord / sigma).astype(int): truncation toward zero of the fp64 quotient
    const long long cx = (long long)((double)x / ss), cy = (long long)((double)y / ss);
    const long long cl = (long long)(Y / sl), cu = (long long)(U / sc), cv = (long long)(V / sc);
    // sum_i coord_i * 255^i  (exact in the reference's fp64 too)
    const long long key = cx + 255LL * (cy + 255LL * (cl + 255LL * (cu + 255LL * cv)));
    // the reference's hash needs every coordinate in [0, 255) (guide * scale in [0, 255]); outside it the hash aliases and --
    // with the 40-bit sort chosen from the sigma bounds -- vertex ids are no longer np.unique's.  Flag it.
    if ((cx | cy | cl | cu | cv) < 0 || cx >= 255 || cy >= 255 || cl >= 255 || cu >= 255 || cv >= 255) at<int>(ws, L.oor)[b] = 1;
    at<unsigned long long>(ws, L.keysA)[t] = (unsigned long long)key;
    at<int>(ws, L.payA)[t] = (int)p;
}

// ---------------------------------------------------------------------------------------------
// 2. stable LSD radix sort of (hash, pixel) on 8-bit digits, all images per launch, then
//    unique -> vertex ids (np.unique, BilateralGrid.py:64-65).
//    Each image is cut into SORT_WPI warp-chunks of consecutive elements; per pass:
//      hist    : every warp counts the digits of its chunk            -> hist[img][digit][warp]
//      scan    : one CTA per image, exclusive scan in (digit, warp) order (stable bucket offsets)
//      scatter : every warp re-reads its chunk in order, ranks equal digits inside each group of
//                32 keys with __match_any_sync and writes (key, pixel) to its private bucket cursor
//    (the first version ran the whole sort in ONE 1024-thread CTA per image: 3.2 ms; profiles/r01)
// ---------------------------------------------------------------------------------------------
constexpr int SORT_WPI = 256;        // warp-chunks per image
constexpr int SORT_CTA_WARPS = 8;    // warps per CTA in hist / scatter / unique kernels

struct SortIO {
    const unsigned long long* kin;
    const int* pin;
    unsigned long long* kout;
    int* pout;
};

__device__ __forceinline__ SortIO sort_io(void* ws, const BgLayout& L, int b, int pass) {
    const size_t o = (size_t)b * L.P;
    unsigned long long* kA = at<unsigned long long>(ws, L.keysA) + o;
    unsigned long long* kB = at<unsigned long long>(ws, L.keysB) + o;
    int* pA = at<int>(ws, L.payA) + o;
    int* pB = at<int>(ws, L.payB) + o;
    return (pass & 1) ? SortIO{kB, pB, kA, pA} : SortIO{kA, pA, kB, pB};
}

__global__ void __launch_bounds__(SORT_CTA_WARPS * 32) bg_sort_hist_kernel(void* ws, BgLayout L, int pass) {
    __shared__ unsigned sh[SORT_CTA_WARPS][256];
    const int b = blockIdx.y, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int warp = blockIdx.x * SORT_CTA_WARPS + wl;  // chunk id inside the image
    const size_t P = L.P, chunk = (P + SORT_WPI - 1) / SORT_WPI;
    const size_t beg = (size_t)warp * chunk, end = beg + chunk < P ? beg + chunk : P;
    const SortIO io = sort_io(ws, L, b, pass);
    const int shift = 8 * pass;
    for (int i = lane; i < 256; i += 32) sh[wl][i] = 0;
    __syncwarp();
    for (size_t i = beg + lane; i < end; i += 32) atomicAdd(&sh[wl][(unsigned)(io.kin[i] >> shift) & 255u], 1u);
    __syncwarp();
    unsigned* hist = at<unsigned>(ws, L.hist) + (size_t)b * 256 * SORT_WPI;
    unsigned* dtot = at<unsigned>(ws, L.dtot) + ((size_t)pass * L.B + b) * 256;
    for (int d = lane; d < 256; d += 32) {
        const unsigned c = sh[wl][d];
        hist[(size_t)d * SORT_WPI + warp] = c;
        if (c) atomicAdd(&dtot[d], c);
    }
}

// exclusive scan of the counters in (digit, chunk) order.  One warp per digit: its base is the sum of the totals of the
// smaller digits (dtot, accumulated by the histogram kernel), its 256 chunk counters are 8 consecutive entries per lane.
__global__ void __launch_bounds__(256) bg_sort_scan_kernel(void* ws, BgLayout L, int pass) {
    static_assert(SORT_WPI == 256, "one lane owns 8 consecutive chunk counters");
    const int b = blockIdx.y, lane = threadIdx.x & 31, d = blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned* hist = at<unsigned>(ws, L.hist) + ((size_t)b * 256 + d) * SORT_WPI;
    const unsigned* dtot = at<unsigned>(ws, L.dtot) + ((size_t)pass * L.B + b) * 256;
    unsigned base = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) base += (lane * 8 + j < d) ? dtot[lane * 8 + j] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    const uint4 lo = reinterpret_cast<const uint4*>(hist)[lane * 2], hi = reinterpret_cast<const uint4*>(hist)[lane * 2 + 1];
    const unsigned loc[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += loc[j];
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    unsigned run = base + inc - s, out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        out[j] = run;
        run += loc[j];
    }
    reinterpret_cast<uint4*>(hist)[lane * 2] = make_uint4(out[0], out[1], out[2], out[3]);
    reinterpret_cast<uint4*>(hist)[lane * 2 + 1] = make_uint4(out[4], out[5], out[6], out[7]);
}

__global__ void __launch_bounds__(SORT_CTA_WARPS * 32) bg_sort_scatter_kernel(void* ws, BgLayout L, int pass) {
    __shared__ unsigned cur[SORT_CTA_WARPS][256];
    const int b = blockIdx.y, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int warp = blockIdx.x * SORT_CTA_WARPS + wl;
    const size_t P = L.P, chunk = (P + SORT_WPI - 1) / SORT_WPI;
    const size_t beg = (size_t)warp * chunk, end = beg + chunk < P ? beg + chunk : P;
    const SortIO io = sort_io(ws, L, b, pass);
    const int shift = 8 * pass;
    const unsigned* hist = at<unsigned>(ws, L.hist) + (size_t)b * 256 * SORT_WPI;
    for (int d = lane; d < 256; d += 32) cur[wl][d] = hist[(size_t)d * SORT_WPI + warp];
    __syncwarp();
    for (size_t i = beg; i < end; i += 32) {
        const bool ok = i + lane < end;
        const unsigned valid = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned long long key = io.kin[i + lane];
            const int pay = io.pin[i + lane];
            const unsigned d = (unsigned)(key >> shift) & 255u;
            const unsigned peers = __match_any_sync(valid, d);
            const unsigned base = cur[wl][d];
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            __syncwarp(valid);
            if (lane == __ffs(peers) - 1) cur[wl][d] = base + __popc(peers);
            io.kout[base + rank] = key;
            io.pout[base + rank] = pay;
        }
        __syncwarp();
    }
}

// unique, phase 1: number of distinct-key heads in every warp-chunk (sorted data in keysB / payB)
__global__ void __launch_bounds__(SORT_CTA_WARPS * 32) bg_unique_count_kernel(void* ws, BgLayout L) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int warp = blockIdx.x * SORT_CTA_WARPS + wl;
    const size_t P = L.P, chunk = (P + SORT_WPI - 1) / SORT_WPI;
    const size_t beg = (size_t)warp * chunk, end = beg + chunk < P ? beg + chunk : P;
    const unsigned long long* k = at<unsigned long long>(ws, L.keysB) + (size_t)b * P;
    unsigned cnt = 0;
    for (size_t i = beg + lane; i < end; i += 32) cnt += (i == 0 || k[i - 1] != k[i]) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0) at<unsigned>(ws, L.hist)[(size_t)b * 256 * SORT_WPI + warp] = cnt;
}

// unique, phase 2: vertex id = rank of the key among the sorted unique keys
__global__ void __launch_bounds__(SORT_CTA_WARPS * 32) bg_unique_rank_kernel(void* ws, BgLayout L) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int warp = blockIdx.x * SORT_CTA_WARPS + wl;
    const size_t P = L.P, chunk = (P + SORT_WPI - 1) / SORT_WPI;
    const size_t beg = (size_t)warp * chunk, end = beg + chunk < P ? beg + chunk : P;
    const unsigned long long* k = at<unsigned long long>(ws, L.keysB) + (size_t)b * P;
    const int* sp = at<int>(ws, L.payB) + (size_t)b * P;
    const unsigned* cnts = at<unsigned>(ws, L.hist) + (size_t)b * 256 * SORT_WPI;
    // heads in the chunks before mine (and, for the last warp, the image total)
    unsigned before = 0, total = 0;
    for (int w = lane; w < SORT_WPI; w += 32) {
        const unsigned c = cnts[w];
        total += c;
        if (w < warp) before += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    int* idx = at<int>(ws, L.idx) + (size_t)b * P;
    int* vstart = at<int>(ws, L.vstart) + (size_t)b * (P + 1);
    unsigned long long* ukeys = at<unsigned long long>(ws, L.ukeys) + (size_t)b * P;
    unsigned run = before;
    for (size_t i = beg; i < end; i += 32) {
        const bool ok = i + lane < end;
        unsigned long long key = 0;
        bool head = false;
        if (ok) {
            key = k[i + lane];
            head = (i + lane == 0) || (k[i + lane - 1] != key);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (ok) {
            const unsigned vid = run + __popc(bal & ((2u << lane) - 1u)) - 1u;  // inclusive count of heads - 1
            if (head) {
                ukeys[vid] = key;
                vstart[vid] = (int)(i + lane);
            }
            idx[sp[i + lane]] = (int)vid;
        }
        run += __popc(bal);
    }
    if (warp == 0 && lane == 0) {
        vstart[total] = (int)P;
        at<int>(ws, L.hdr)[b] = (int)total;
    }
}

// ---------------------------------------------------------------------------------------------
// 3. neighbour table (BilateralGrid.py:72-84): hash +- 255^d located by binary search
// ---------------------------------------------------------------------------------------------
__global__ void bg_nbr_kernel(void* ws, BgLayout L) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t P = L.P;
    const int b = blockIdx.y;
    const int V = at<int>(ws, L.hdr)[b];
    if (t >= (size_t)V) return;
    const unsigned long long* uk = at<unsigned long long>(ws, L.ukeys) + (size_t)b * P;
    int* nbr = at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P;
    const long long key = (long long)uk[t];
    // ten lower-bound searches (hash +- 255^d) advanced in lock-step with the branch-free uniform binary search: every step
    // issues ten independent loads instead of one (the one-search-at-a-time loop was a chain of ~180 dependent L2 loads per
    // vertex: 182 us per build at config 4)
    long long target[2 * BG_DIM];
    int base[2 * BG_DIM];
    {
        long long hv = 1;
#pragma unroll
        for (int d = 0; d < BG_DIM; ++d) {
            target[2 * d] = key - hv;
            target[2 * d + 1] = key + hv;
            base[2 * d] = base[2 * d + 1] = 0;
            hv *= 255;
        }
    }
    for (int n = V; n > 1;) {
        const int half = n >> 1;
#pragma unroll
        for (int j = 0; j < 2 * BG_DIM; ++j) base[j] = ((long long)uk[base[j] + half - 1] < target[j]) ? base[j] + half : base[j];
        n -= half;
    }
#pragma unroll
    for (int j = 0; j < 2 * BG_DIM; ++j) {
        const long long k0 = (long long)uk[base[j]];
        int lo = base[j];                                      // the lower bound is base or base + 1
        bool hit = k0 == target[j];
        if (k0 < target[j]) {
            lo = base[j] + 1;
            hit = lo < V && (long long)uk[lo] == target[j];
        }
        nbr[(size_t)j * P + t] = hit ? lo : -1;
    }
    // pixel counts m0 = S 1
    const int* vs = at<int>(ws, L.vstart) + (size_t)b * (P + 1);
    at<double>(ws, L.m0)[(size_t)b * P + t] = (double)(vs[t + 1] - vs[t]);
    at<double>(ws, L.n0)[(size_t)b * P + t] = 1.0;
}

// blur(x)_v = 2*dim*x_v + sum_dims (x[lo] + x[hi])  in the reference's summation order
// (BilateralGrid.py:96-102: out = 10 x; for each dim: out += B_dim x, B_dim row = [lo, hi] by column)
__device__ __forceinline__ double bg_blur(const double* __restrict__ x, const int* __restrict__ nbr, size_t P, size_t v) {
    double out = 2.0 * BG_DIM * x[v];
#pragma unroll
    for (int d = 0; d < BG_DIM; ++d) {
        const int lo = nbr[(size_t)(2 * d) * P + v], hi = nbr[(size_t)(2 * d + 1) * P + v];
        double t = 0.0;
        bool any = false;
        if (lo >= 0) { t = x[lo]; any = true; }
        if (hi >= 0) { t = any ? t + x[hi] : x[hi]; any = true; }
        if (any) out += t;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// 4. bistochastisation (BilateralGrid.py:109-120)
// ---------------------------------------------------------------------------------------------
__global__ void bg_bisto_iter_kernel(void* ws, BgLayout L, int src) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const size_t P = L.P;
    if (t >= (size_t)at<int>(ws, L.hdr)[b]) return;
    const double* n = at<double>(ws, src ? L.n1 : L.n0) + (size_t)b * P;
    double* nn = at<double>(ws, src ? L.n0 : L.n1) + (size_t)b * P;
    const int* nbr = at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P;
    const double m0 = at<double>(ws, L.m0)[(size_t)b * P + t];
    nn[t] = sqrt(n[t] * m0 / bg_blur(n, nbr, P, t));
}
__global__ void bg_bisto_final_kernel(void* ws, BgLayout L, int src) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const size_t P = L.P;
    if (t >= (size_t)at<int>(ws, L.hdr)[b]) return;
    const double* n = at<double>(ws, src ? L.n1 : L.n0) + (size_t)b * P;
    const int* nbr = at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P;
    at<double>(ws, L.m)[(size_t)b * P + t] = n[t] * bg_blur(n, nbr, P, t);
    if (src) at<double>(ws, L.n0)[(size_t)b * P + t] = n[t];  // final n always in n0
}

// ---------------------------------------------------------------------------------------------
// 5. solver
// ---------------------------------------------------------------------------------------------
#ifndef IRB_BG_RECORDS
#define IRB_BG_RECORDS 1
#endif
struct SolveArgs {
    const float* pred;   // [B,ch,P]  (forward: target; backward: target for grad_conf)
    const float* conf;   // [B,1,P]
    const float* rhs;    // forward: pred ; backward: g_out   [B,ch,P]
    double* x;           // [B,ch,P] solution (yhat buffer of the caller)
    double lam, dmin, tol;
    int adjoint;         // 0: b = S(x*w), y0 = b/max(Sw,1e-10) ; 1: b = S(g), y0 = b/S(1)
    int b0;              // first image of the group this launch works on (grid.y = images in the group)
};

__device__ __forceinline__ double* scal_ptr(void* ws, const BgLayout& L, int b, int c, int it) {
    return at<double>(ws, L.scal) + (((size_t)b * BG_MAX_CH + c) * (BG_MAX_ITERS + 2) + it) * 4;
}
// slot layout: [it] -> {rho, pq, rr, -};   slot BG_MAX_ITERS+1 -> {bb, -, -, -}

__device__ __forceinline__ void block_atomic_add(double v, double* dst) {
    // warp shuffle reduction, then one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(dst, v);
}

// several sums at once: warp shuffle, then shared memory across the block's warps, then ONE atomic per sum and block
template <int N>
__device__ __forceinline__ void block_atomic_add_n(double (&v)[N], double* const (&dst)[N]) {
    __shared__ double part[8][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int i = 0; i < N; ++i) part[w][i] = v[i];
    __syncthreads();
    if (threadIdx.x < N) {
        double t = 0.0;
        for (int j = 0; j < nw; ++j) t += part[j][threadIdx.x];
        if (t != 0.0 && dst[threadIdx.x]) atomicAdd(dst[threadIdx.x], t);
    }
}

// splat (ordered segmented sums), system diagonal, flat initialisation (BilateralGrid.py:134-145 / 160-168)
__global__ void bg_setup_kernel(void* ws, BgLayout L, SolveArgs a) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    const int V = at<int>(ws, L.hdr)[b];
    const bool ok = t < (size_t)V;
    double bsum[BG_MAX_CH] = {0, 0, 0, 0};
    if (ok) {
        const int* vs = at<int>(ws, L.vstart) + (size_t)b * (P + 1);
        const int* sp = at<int>(ws, L.payB) + (size_t)b * P;
        const int s0 = vs[t], s1 = vs[t + 1];
        double wsum = 0.0;
        for (int i = s0; i < s1; ++i) {
            const int p = sp[i];
            const double w = (double)a.conf[(size_t)b * P + p];
            wsum += w;
            for (int c = 0; c < L.ch; ++c) {
                const double xv = (double)a.rhs[((size_t)b * L.ch + c) * P + p];
                bsum[c] += a.adjoint ? xv : xv * w;
            }
        }
        const double n = at<double>(ws, L.n0)[(size_t)b * P + t], m = at<double>(ws, L.m)[(size_t)b * P + t];
        const double diag = a.lam * (m - n * (2.0 * BG_DIM) * n) + wsum;
        at<double>(ws, L.wsplat)[(size_t)b * P + t] = wsum;
        at<double>(ws, L.minv)[(size_t)b * P + t] = 1.0 / fmax(diag, a.dmin);
        const double den = a.adjoint ? at<double>(ws, L.m0)[(size_t)b * P + t] : fmax(wsum, 1e-10);
        double x0[BG_MAX_CH] = {0, 0, 0, 0};
        for (int c = 0; c < L.ch; ++c) {
            at<double>(ws, L.bvec)[((size_t)b * L.ch + c) * P + t] = bsum[c];
            x0[c] = bsum[c] / den;
            a.x[((size_t)b * L.ch + c) * P + t] = x0[c];
        }
#if IRB_BG_RECORDS
        // vertex-major record {x0_0..x0_2, n} for bg_r0_rec_kernel's neighbour gathers (same slot the PCG records use later)
        if (L.ch == 1) at<double2>(ws, L.np)[(size_t)b * P + t] = make_double2(x0[0], n);
        else if (L.ch <= 3) at<double4>(ws, L.np)[(size_t)b * P + t] = make_double4(x0[0], x0[1], x0[2], n);
#endif
    }
    for (int c = 0; c < L.ch; ++c) block_atomic_add(ok ? bsum[c] * bsum[c] : 0.0, scal_ptr(ws, L, b, c, BG_MAX_ITERS + 1));
}

// A y = lam (m y - n blur(n y)) + wsplat y, with y given by a functor over vertex ids
template <typename F>
__device__ __forceinline__ double bg_matvec(F y, const double* __restrict__ n, const int* __restrict__ nbr, size_t P, size_t v,
                                            double m, double wsp, double lam) {
    double out = 2.0 * BG_DIM * (n[v] * y(v));
#pragma unroll
    for (int d = 0; d < BG_DIM; ++d) {
        const int lo = nbr[(size_t)(2 * d) * P + v], hi = nbr[(size_t)(2 * d + 1) * P + v];
        double t = 0.0;
        bool any = false;
        if (lo >= 0) { t = n[lo] * y(lo); any = true; }
        if (hi >= 0) { const double u = n[hi] * y(hi); t = any ? t + u : u; any = true; }
        if (any) out += t;
    }
    const double yv = y(v);
    return lam * (m * yv - n[v] * out) + wsp * yv;
}

// r0 = b - A x0 ; rho0 = <r0, M r0> ; rr0 = <r0, r0>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_r0_kernel(void* ws, BgLayout L, SolveArgs a) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    const int V = at<int>(ws, L.hdr)[b];
    const bool ok = t < (size_t)V;
    const double* n = at<double>(ws, L.n0) + (size_t)b * P;
    const int* nbr = at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P;
    for (int c = 0; c < L.ch; ++c) {
        double rho = 0.0, rr = 0.0;
        if (ok) {
            const double* x = a.x + ((size_t)b * L.ch + c) * P;
            const double Ax = bg_matvec([&](size_t j) { return x[j]; }, n, nbr, P, t, at<double>(ws, L.m)[(size_t)b * P + t],
                                        at<double>(ws, L.wsplat)[(size_t)b * P + t], a.lam);
            const double r = at<double>(ws, L.bvec)[((size_t)b * L.ch + c) * P + t] - Ax;
            at<double>(ws, L.r)[((size_t)b * L.ch + c) * P + t] = r;
            rho = r * at<double>(ws, L.minv)[(size_t)b * P + t] * r;
            rr = r * r;
        }
        double* s = scal_ptr(ws, L, b, c, 0);
        block_atomic_add(rho, s + 0);
        block_atomic_add(rr, s + 2);
    }
}

// Cache policy of the PCG kernels: everything but `np` is streamed once per launch (52 doubles per vertex and iteration,
// 5x the L2), `np` (32 B per vertex, 50 MB at config 4) is gathered ten times per vertex by bg_q.  Streaming accesses use
// the evict-first policy (ld/st.global.cs) so that they do not push `np` out between bg_p (writes it) and bg_q.
#ifndef IRB_BG_STREAM_HINTS
#define IRB_BG_STREAM_HINTS 1
#endif
template <typename T> __device__ __forceinline__ T ld_stream(const T* p) {
#if IRB_BG_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename T> __device__ __forceinline__ void st_stream(T* p, T v) {
#if IRB_BG_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}

__device__ __forceinline__ bool bg_done(void* ws, const BgLayout& L, int b, int c, int it, double tol) {
    const double* s = scal_ptr(ws, L, b, c, it);
    const double bb = scal_ptr(ws, L, b, c, BG_MAX_ITERS + 1)[0];
    // scipy cg: stop when ||r|| < max(atol, rtol*||b||); ||b|| == 0 returns immediately
    return bb == 0.0 || s[2] < tol * tol * bb || s[0] == 0.0;
}

// iteration `it`, step 1:  p = z + (rho/rho_prev) p   (z = M r);  also np[v][c] = n_v p_v[c], interleaved over
// channels so that the mat-vec's neighbour gathers fetch all channels of a vertex in ONE 32-byte sector
template <int CH>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_p_kernel(void* ws, BgLayout L, SolveArgs a, int it) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    if (t >= (size_t)at<int>(ws, L.hdr)[b]) return;
    bool done[CH];
    double beta[CH], rv[CH], pold[CH];
    const double* pprev = at<double>(ws, (it & 1) ? L.p1 : L.p0);
    double* pnext = at<double>(ws, (it & 1) ? L.p0 : L.p1);
#pragma unroll
    for (int c = 0; c < CH; ++c) {   // all loads of the thread are issued before the first use
        done[c] = bg_done(ws, L, b, c, it, a.tol);
        beta[c] = (it > 0 && !done[c]) ? scal_ptr(ws, L, b, c, it)[0] / scal_ptr(ws, L, b, c, it - 1)[0] : 0.0;
        const size_t o = ((size_t)b * CH + c) * P + t;
        rv[c] = done[c] ? 0.0 : ld_stream(at<double>(ws, L.r) + o);
        pold[c] = (it > 0 && !done[c]) ? ld_stream(pprev + o) : 0.0;
    }
    const double nv = ld_stream(at<double>(ws, L.n0) + (size_t)b * P + t);
    const double mi = ld_stream(at<double>(ws, L.minv) + (size_t)b * P + t);
    double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        if (done[c]) continue;
        const double z = mi * rv[c];
        const double pv = (it > 0) ? fma(beta[c], pold[c], z) : z;
        st_stream(pnext + ((size_t)b * CH + c) * P + t, pv);
        out[c] = nv * pv;
    }
    double4* np = at<double4>(ws, L.np) + (size_t)b * P + t;
    *np = make_double4(out[0], out[1], out[2], out[3]);
}

// iteration `it`, step 2:  q = A p = lam (m p - n blur(n p)) + wsplat p ;  pq = <p,q>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_q_kernel(void* ws, BgLayout L, SolveArgs a, int it) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    const int V = at<int>(ws, L.hdr)[b];
    const bool ok = t < (size_t)V;
    double bl[4] = {0.0, 0.0, 0.0, 0.0};
    double nv = 0.0, mv = 0.0, wsp = 0.0;
    if (ok) {
        const double4* np = at<double4>(ws, L.np) + (size_t)b * P;
        const int* nbr = at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P;
        const double4 own = np[t];
        bl[0] = 2.0 * BG_DIM * own.x; bl[1] = 2.0 * BG_DIM * own.y; bl[2] = 2.0 * BG_DIM * own.z; bl[3] = 2.0 * BG_DIM * own.w;
#pragma unroll
        for (int d = 0; d < BG_DIM; ++d) {   // same summation order as bg_blur: per dimension (lo + hi), then accumulate
            const int lo = ld_stream(nbr + (size_t)(2 * d) * P + t), hi = ld_stream(nbr + (size_t)(2 * d + 1) * P + t);
            double4 tt = make_double4(0.0, 0.0, 0.0, 0.0);
            bool any = false;
            if (lo >= 0) { tt = np[lo]; any = true; }
            if (hi >= 0) {
                const double4 u = np[hi];
                if (any) { tt.x += u.x; tt.y += u.y; tt.z += u.z; tt.w += u.w; } else { tt = u; }
                any = true;
            }
            if (any) { bl[0] += tt.x; bl[1] += tt.y; bl[2] += tt.z; bl[3] += tt.w; }
        }
        nv = ld_stream(at<double>(ws, L.n0) + (size_t)b * P + t);
        mv = ld_stream(at<double>(ws, L.m) + (size_t)b * P + t);
        wsp = ld_stream(at<double>(ws, L.wsplat) + (size_t)b * P + t);
    }
    double pq[BG_MAX_CH] = {0.0, 0.0, 0.0, 0.0};
    double* dst[BG_MAX_CH] = {nullptr, nullptr, nullptr, nullptr};
    for (int c = 0; c < L.ch; ++c) {
        if (bg_done(ws, L, b, c, it, a.tol)) continue;  // uniform per (image, channel)
        dst[c] = scal_ptr(ws, L, b, c, it) + 1;
        if (ok) {
            const size_t o = ((size_t)b * L.ch + c) * P + t;
            const double pv = ld_stream(at<double>(ws, (it & 1) ? L.p0 : L.p1) + o);
            const double q = a.lam * (mv * pv - nv * bl[c]) + wsp * pv;
            st_stream(at<double>(ws, L.q) + o, q);
            pq[c] = pv * q;
        }
    }
    block_atomic_add_n<BG_MAX_CH>(pq, dst);
}

// iteration `it`, second half: alpha = rho/pq ; x += alpha p ; r -= alpha q ; next rho, rr
template <int CH>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_xr_kernel(void* ws, BgLayout L, SolveArgs a, int it) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    const int V = at<int>(ws, L.hdr)[b];
    const bool ok = t < (size_t)V;
    double sums[2 * CH];
    double* dst[2 * CH];
    double alpha[CH], pv[CH], xv[CH], rv[CH], qv[CH];
    const double* pcur = at<double>(ws, (it & 1) ? L.p0 : L.p1);
#pragma unroll
    for (int c = 0; c < CH; ++c) {   // all loads first, then the arithmetic, then one reduction for the 2*CH sums
        const bool done = bg_done(ws, L, b, c, it, a.tol);
        const double* s = scal_ptr(ws, L, b, c, it);
        double* sn = scal_ptr(ws, L, b, c, it + 1);
        dst[2 * c] = done ? nullptr : sn + 0;
        dst[2 * c + 1] = done ? nullptr : sn + 2;
        alpha[c] = done ? 0.0 : s[0] / s[1];
        const size_t o = ((size_t)b * CH + c) * P + t;
        const bool ld = ok && !done;
        pv[c] = ld ? ld_stream(pcur + o) : 0.0;
        xv[c] = ld ? ld_stream(a.x + o) : 0.0;
        rv[c] = ld ? ld_stream(at<double>(ws, L.r) + o) : 0.0;
        qv[c] = ld ? ld_stream(at<double>(ws, L.q) + o) : 0.0;
    }
    const double mi = ok ? ld_stream(at<double>(ws, L.minv) + (size_t)b * P + t) : 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        sums[2 * c] = sums[2 * c + 1] = 0.0;
        if (ok && dst[2 * c]) {
            const size_t o = ((size_t)b * CH + c) * P + t;
            st_stream(a.x + o, xv[c] + alpha[c] * pv[c]);
            const double r = rv[c] - alpha[c] * qv[c];
            st_stream(at<double>(ws, L.r) + o, r);
            sums[2 * c] = r * mi * r;
            sums[2 * c + 1] = r * r;
        }
    }
    block_atomic_add_n<2 * CH>(sums, dst);
}

// ---- record form of the three PCG kernels (1..3 channels) -------------------------------------------------------------
// The legacy kernels above keep the search direction twice: p (channel-major, ping-pong) and np = n * p (vertex-major, what
// the neighbour gathers read).  Here ONE vertex-major record {p_0..p_{CH-1}, n} (double2 for one channel, double4 for 2-3)
// replaces both: bg_q multiplies n_j * p_j after the gather -- the same rounded product that bg_p used to store (__dmul_rn, not
// contracted), so the arithmetic of the solve is unchanged (results still vary in the last bits with the order of the fp64
// atomics of the dot products, as before) -- and the own p and n come out of the same load.  Doubles moved per vertex
// and iteration: 3 channels 52 -> 46, 1 channel 30 -> 24; a one-channel gather fetches 16 instead of 32 bytes.
template <int CH> struct BgRec { typedef double4 T; };
template <> struct BgRec<1> { typedef double2 T; };
template <int CH> __device__ __forceinline__ void rec_unpack(const typename BgRec<CH>::T& r, double (&p)[3], double& n);
template <> __device__ __forceinline__ void rec_unpack<1>(const double2& r, double (&p)[3], double& n) { p[0] = r.x; p[1] = p[2] = 0.0; n = r.y; }
template <> __device__ __forceinline__ void rec_unpack<2>(const double4& r, double (&p)[3], double& n) { p[0] = r.x; p[1] = r.y; p[2] = 0.0; n = r.w; }
template <> __device__ __forceinline__ void rec_unpack<3>(const double4& r, double (&p)[3], double& n) { p[0] = r.x; p[1] = r.y; p[2] = r.z; n = r.w; }
template <int CH> __device__ __forceinline__ typename BgRec<CH>::T rec_pack(const double (&p)[3], double n);
template <> __device__ __forceinline__ double2 rec_pack<1>(const double (&p)[3], double n) { return make_double2(p[0], n); }
template <> __device__ __forceinline__ double4 rec_pack<2>(const double (&p)[3], double n) { return make_double4(p[0], p[1], 0.0, n); }
template <> __device__ __forceinline__ double4 rec_pack<3>(const double (&p)[3], double n) { return make_double4(p[0], p[1], p[2], n); }

template <int CH>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_p_rec_kernel(void* ws, BgLayout L, SolveArgs a, int it) {
    typedef typename BgRec<CH>::T Rec;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    if (t >= (size_t)at<int>(ws, L.hdr)[b]) return;
    Rec* rec = at<Rec>(ws, L.np) + (size_t)b * P + t;
    bool done[CH];
    double beta[CH], rv[CH], pold[3] = {0.0, 0.0, 0.0}, nv;
    if (it > 0) rec_unpack<CH>(*rec, pold, nv);
    else nv = ld_stream(at<double>(ws, L.n0) + (size_t)b * P + t);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        done[c] = bg_done(ws, L, b, c, it, a.tol);
        beta[c] = (it > 0 && !done[c]) ? scal_ptr(ws, L, b, c, it)[0] / scal_ptr(ws, L, b, c, it - 1)[0] : 0.0;
        rv[c] = done[c] ? 0.0 : ld_stream(at<double>(ws, L.r) + ((size_t)b * CH + c) * P + t);
    }
    const double mi = ld_stream(at<double>(ws, L.minv) + (size_t)b * P + t);
    double pn[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        if (done[c]) continue;   // a finished channel's slot reads as 0, nobody uses it
        const double z = mi * rv[c];
        pn[c] = (it > 0) ? fma(beta[c], pold[c], z) : z;
    }
    *rec = rec_pack<CH>(pn, nv);
}

// blur part of A y for the vertex t from the records: bl[c] = 2 dim n_t y_t[c] + sum_dims (n_lo y_lo[c] + n_hi y_hi[c]), the
// summation order of bg_blur; also returns the vertex's own y and n
template <int CH>
__device__ __forceinline__ void rec_blur(const typename BgRec<CH>::T* __restrict__ rec, const int* __restrict__ nbr, size_t P,
                                         size_t t, double (&bl)[3], double (&yv)[3], double& nv) {
    rec_unpack<CH>(rec[t], yv, nv);
#pragma unroll
    for (int c = 0; c < 3; ++c) bl[c] = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) bl[c] = 2.0 * BG_DIM * __dmul_rn(nv, yv[c]);
#pragma unroll
    for (int d = 0; d < BG_DIM; ++d) {
        const int lo = ld_stream(nbr + (size_t)(2 * d) * P + t), hi = ld_stream(nbr + (size_t)(2 * d + 1) * P + t);
        double tt[3] = {0.0, 0.0, 0.0};
        bool any = false;
        if (lo >= 0) {
            double pj[3], nj;
            rec_unpack<CH>(rec[lo], pj, nj);
#pragma unroll
            for (int c = 0; c < CH; ++c) tt[c] = __dmul_rn(nj, pj[c]);   // rounded product, never contracted into the add below
            any = true;
        }
        if (hi >= 0) {
            double pj[3], nj;
            rec_unpack<CH>(rec[hi], pj, nj);
#pragma unroll
            for (int c = 0; c < CH; ++c) tt[c] = any ? tt[c] + __dmul_rn(nj, pj[c]) : __dmul_rn(nj, pj[c]);
            any = true;
        }
        if (any) {
#pragma unroll
            for (int c = 0; c < CH; ++c) bl[c] += tt[c];
        }
    }
}

template <int CH>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_q_rec_kernel(void* ws, BgLayout L, SolveArgs a, int it) {
    typedef typename BgRec<CH>::T Rec;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    const int V = at<int>(ws, L.hdr)[b];
    const bool ok = t < (size_t)V;
    double bl[3] = {0.0, 0.0, 0.0}, pv[3] = {0.0, 0.0, 0.0};
    double nv = 0.0, mv = 0.0, wsp = 0.0;
    if (ok) {
        rec_blur<CH>(at<Rec>(ws, L.np) + (size_t)b * P, at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P, P, t, bl, pv, nv);
        mv = ld_stream(at<double>(ws, L.m) + (size_t)b * P + t);
        wsp = ld_stream(at<double>(ws, L.wsplat) + (size_t)b * P + t);
    }
    double pq[BG_MAX_CH] = {0.0, 0.0, 0.0, 0.0};
    double* dst[BG_MAX_CH] = {nullptr, nullptr, nullptr, nullptr};
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        if (bg_done(ws, L, b, c, it, a.tol)) continue;  // uniform per (image, channel)
        dst[c] = scal_ptr(ws, L, b, c, it) + 1;
        if (ok) {
            const double q = a.lam * (mv * pv[c] - nv * bl[c]) + wsp * pv[c];
            st_stream(at<double>(ws, L.q) + ((size_t)b * CH + c) * P + t, q);
            pq[c] = pv[c] * q;
        }
    }
    block_atomic_add_n<BG_MAX_CH>(pq, dst);
}

// r0 = b - A x0 ; rho0 = <r0, M r0> ; rr0 = <r0, r0>, x0 read from the records bg_setup wrote (one 32 / 16-byte gather per
// neighbour for all channels; the legacy bg_r0_kernel gathers n and x separately, per channel: 60 gathers per vertex)
template <int CH>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_r0_rec_kernel(void* ws, BgLayout L, SolveArgs a) {
    typedef typename BgRec<CH>::T Rec;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    const int V = at<int>(ws, L.hdr)[b];
    const bool ok = t < (size_t)V;
    double sums[2 * CH];
    double* dst[2 * CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        sums[2 * c] = sums[2 * c + 1] = 0.0;
        dst[2 * c] = scal_ptr(ws, L, b, c, 0) + 0;
        dst[2 * c + 1] = scal_ptr(ws, L, b, c, 0) + 2;
    }
    if (ok) {
        double bl[3], xv[3], nv;
        rec_blur<CH>(at<Rec>(ws, L.np) + (size_t)b * P, at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P, P, t, bl, xv, nv);
        const double mv = ld_stream(at<double>(ws, L.m) + (size_t)b * P + t);
        const double wsp = ld_stream(at<double>(ws, L.wsplat) + (size_t)b * P + t);
        const double mi = ld_stream(at<double>(ws, L.minv) + (size_t)b * P + t);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const size_t o = ((size_t)b * CH + c) * P + t;
            const double Ax = a.lam * (mv * xv[c] - nv * bl[c]) + wsp * xv[c];
            const double r = ld_stream(at<double>(ws, L.bvec) + o) - Ax;
            st_stream(at<double>(ws, L.r) + o, r);
            sums[2 * c] = r * mi * r;
            sums[2 * c + 1] = r * r;
        }
    }
    block_atomic_add_n<2 * CH>(sums, dst);
}

template <int CH>
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_xr_rec_kernel(void* ws, BgLayout L, SolveArgs a, int it) {
    typedef typename BgRec<CH>::T Rec;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y + a.b0;
    const size_t P = L.P;
    const int V = at<int>(ws, L.hdr)[b];
    const bool ok = t < (size_t)V;
    double sums[2 * CH];
    double* dst[2 * CH];
    double alpha[CH], pv[3] = {0.0, 0.0, 0.0}, xv[CH], rv[CH], qv[CH], nv;
    if (ok) rec_unpack<CH>(at<Rec>(ws, L.np)[(size_t)b * P + t], pv, nv);
#pragma unroll
    for (int c = 0; c < CH; ++c) {   // all loads first, then the arithmetic, then one reduction for the 2*CH sums
        const bool done = bg_done(ws, L, b, c, it, a.tol);
        const double* s = scal_ptr(ws, L, b, c, it);
        double* sn = scal_ptr(ws, L, b, c, it + 1);
        dst[2 * c] = done ? nullptr : sn + 0;
        dst[2 * c + 1] = done ? nullptr : sn + 2;
        alpha[c] = done ? 0.0 : s[0] / s[1];
        const size_t o = ((size_t)b * CH + c) * P + t;
        const bool ld = ok && !done;
        xv[c] = ld ? ld_stream(a.x + o) : 0.0;
        rv[c] = ld ? ld_stream(at<double>(ws, L.r) + o) : 0.0;
        qv[c] = ld ? ld_stream(at<double>(ws, L.q) + o) : 0.0;
    }
    const double mi = ok ? ld_stream(at<double>(ws, L.minv) + (size_t)b * P + t) : 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        sums[2 * c] = sums[2 * c + 1] = 0.0;
        if (ok && dst[2 * c]) {
            const size_t o = ((size_t)b * CH + c) * P + t;
            st_stream(a.x + o, xv[c] + alpha[c] * pv[c]);
            const double r = rv[c] - alpha[c] * qv[c];
            st_stream(at<double>(ws, L.r) + o, r);
            sums[2 * c] = r * mi * r;
            sums[2 * c + 1] = r * r;
        }
    }
    block_atomic_add_n<2 * CH>(sums, dst);
}

// slice (BilateralGrid.py:93-94) -> fp32 output
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_slice_kernel(const void* ws, BgLayout L, const double* __restrict__ x, float* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (t >= L.P) return;
    const int v = at<int>(ws, L.idx)[(size_t)b * L.P + t];
    for (int c = 0; c < L.ch; ++c) out[((size_t)b * L.ch + c) * L.P + t] = (float)x[((size_t)b * L.ch + c) * L.P + v];
}

// backward epilogue (BilateralGrid.py:174-183, 203-205)
__global__ void __launch_bounds__(256, IRB_BG_MINB) bg_grad_kernel(const void* ws, BgLayout L, const double* __restrict__ gh, const double* __restrict__ yhat,
                               const float* __restrict__ pred, const float* __restrict__ conf, float* __restrict__ g_pred,
                               float* __restrict__ g_conf) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (t >= L.P) return;
    const int v = at<int>(ws, L.idx)[(size_t)b * L.P + t];
    const double w = (double)conf[(size_t)b * L.P + t];
    double gc = 0.0;
    for (int c = 0; c < L.ch; ++c) {
        const size_t ov = ((size_t)b * L.ch + c) * L.P + v, op = ((size_t)b * L.ch + c) * L.P + t;
        const double g = gh[ov];
        if (g_pred) g_pred[op] = (float)(g * w);
        gc += -1.0 * (g * yhat[ov]) + g * (double)pred[op];
    }
    if (g_conf) g_conf[(size_t)b * L.P + t] = (float)gc;
}

// Images per PCG launch.  Measured on B200 (profiles/r02, config 4, fwd+bwd): groups of 1 / 2 / 3 / 4 / 8 images = 9.25 / 7.66 /
// 7.40 / 7.17 / 7.03 ms -- the kernels are latency-bound (ncu: 55-85 % long-scoreboard stalls at 1.1-2.7 TB/s), so keeping a
// group's vectors L2-resident does not pay and the largest launch wins: default = the whole batch.  IRB_BG_GROUP overrides.
static int pcg_group(const BgLayout& L) {
    if (const char* g = getenv("IRB_BG_GROUP")) {
        const int v = atoi(g);
        if (v > 0) return v < L.B ? v : L.B;
    }
    return L.B;
}

static int run_pcg(void* ws, const BgLayout& L, SolveArgs a, int maxiter, cudaStream_t st) {
    const unsigned gx = (unsigned)((L.P + 255) / 256);
    cudaMemsetAsync(at<double>(ws, L.scal), 0, 8 * (size_t)L.B * BG_MAX_CH * (BG_MAX_ITERS + 2) * 4, st);
    const int G = pcg_group(L);
    int launches = 0;
    for (int b0 = 0; b0 < L.B; b0 += G) {
        const dim3 grid(gx, (unsigned)(L.B - b0 < G ? L.B - b0 : G));
        a.b0 = b0;
        bg_setup_kernel<<<grid, 256, 0, st>>>(ws, L, a);
#if IRB_BG_RECORDS
        if (L.ch == 1) bg_r0_rec_kernel<1><<<grid, 256, 0, st>>>(ws, L, a);
        else if (L.ch == 2) bg_r0_rec_kernel<2><<<grid, 256, 0, st>>>(ws, L, a);
        else if (L.ch == 3) bg_r0_rec_kernel<3><<<grid, 256, 0, st>>>(ws, L, a);
        else
#endif
        bg_r0_kernel<<<grid, 256, 0, st>>>(ws, L, a);
        for (int it = 0; it < maxiter; ++it) {
#if IRB_BG_RECORDS
            if (L.ch <= 3) {
                switch (L.ch) {
                    case 1:
                        bg_p_rec_kernel<1><<<grid, 256, 0, st>>>(ws, L, a, it);
                        bg_q_rec_kernel<1><<<grid, 256, 0, st>>>(ws, L, a, it);
                        bg_xr_rec_kernel<1><<<grid, 256, 0, st>>>(ws, L, a, it);
                        break;
                    case 2:
                        bg_p_rec_kernel<2><<<grid, 256, 0, st>>>(ws, L, a, it);
                        bg_q_rec_kernel<2><<<grid, 256, 0, st>>>(ws, L, a, it);
                        bg_xr_rec_kernel<2><<<grid, 256, 0, st>>>(ws, L, a, it);
                        break;
                    default:
                        bg_p_rec_kernel<3><<<grid, 256, 0, st>>>(ws, L, a, it);
                        bg_q_rec_kernel<3><<<grid, 256, 0, st>>>(ws, L, a, it);
                        bg_xr_rec_kernel<3><<<grid, 256, 0, st>>>(ws, L, a, it);
                        break;
                }
                continue;
            }
#endif
            switch (L.ch) {
                case 1: bg_p_kernel<1><<<grid, 256, 0, st>>>(ws, L, a, it); break;
                case 2: bg_p_kernel<2><<<grid, 256, 0, st>>>(ws, L, a, it); break;
                case 3: bg_p_kernel<3><<<grid, 256, 0, st>>>(ws, L, a, it); break;
                default: bg_p_kernel<4><<<grid, 256, 0, st>>>(ws, L, a, it); break;
            }
            bg_q_kernel<<<grid, 256, 0, st>>>(ws, L, a, it);
            switch (L.ch) {
                case 1: bg_xr_kernel<1><<<grid, 256, 0, st>>>(ws, L, a, it); break;
                case 2: bg_xr_kernel<2><<<grid, 256, 0, st>>>(ws, L, a, it); break;
                case 3: bg_xr_kernel<3><<<grid, 256, 0, st>>>(ws, L, a, it); break;
                default: bg_xr_kernel<4><<<grid, 256, 0, st>>>(ws, L, a, it); break;
            }
        }
        launches += 2 + 3 * maxiter;
    }
    return check_launch("bilateral pcg", launches);
}

}  // namespace irb

using namespace irb;

extern "C" size_t irb_bg_workspace_bytes(int B, int H, int W, int ch) {
    if (B <= 0 || H <= 0 || W <= 0 || ch <= 0 || ch > BG_MAX_CH) return 0;
    return make_layout(B, H, W, ch).total;
}

// ---------------------------------------------------------------------------------------------
// 6. primitive operators of the grid facade (BilateralGrid.py:90-107): splat = S x (ordered segmented sums, the CSR
//    product's summation order), slice = S^T y, blur = 2 dim y + sum_dims B_dim y.  fp64 in / out like the reference.
//    op 0 splat: x [B,ch,P] per pixel -> out [B,ch,P] per vertex (entries >= V zeroed); 1 slice; 2 blur (vertex -> vertex)
// ---------------------------------------------------------------------------------------------
__global__ void bg_op_kernel(const void* ws, BgLayout L, int op, const double* __restrict__ x, double* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const size_t P = L.P;
    if (t >= P) return;
    const int V = at<int>(ws, L.hdr)[b];
    for (int c = 0; c < L.ch; ++c) {
        const double* xc = x + ((size_t)b * L.ch + c) * P;
        double* oc = out + ((size_t)b * L.ch + c) * P;
        if (op == 1) {
            oc[t] = xc[at<int>(ws, L.idx)[(size_t)b * P + t]];
        } else if (t >= (size_t)V) {
            oc[t] = 0.0;
        } else if (op == 0) {
            const int* vs = at<int>(ws, L.vstart) + (size_t)b * (P + 1);
            const int* sp = at<int>(ws, L.payB) + (size_t)b * P;
            double acc = 0.0;
            for (int i = vs[t]; i < vs[t + 1]; ++i) acc += xc[sp[i]];
            oc[t] = acc;
        } else {
            oc[t] = bg_blur(xc, at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P, P, t);
        }
    }
}

// directed neighbour links per image (E of SURVEY 8d)
__global__ void bg_links_kernel(const void* ws, BgLayout L, unsigned long long* __restrict__ links) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const size_t P = L.P;
    int cnt = 0;
    if (t < (size_t)at<int>(ws, L.hdr)[b]) {
        const int* nbr = at<int>(ws, L.nbr) + (size_t)b * 2 * BG_DIM * P;
        for (int d = 0; d < 2 * BG_DIM; ++d) cnt += nbr[(size_t)d * P + t] >= 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(links + b, (unsigned long long)cnt);
}

static int bg_check(const void* ws, int B, int H, int W, int ch, const char* who) {
    IRB_REQUIRE(ws != nullptr, "%s: null workspace", who);
    IRB_REQUIRE(B > 0 && H > 0 && W > 0, "%s: non-positive shape", who);
    IRB_REQUIRE(ch > 0 && ch <= BG_MAX_CH, "%s: channels must be 1..%d (got %d)", who, BG_MAX_CH, ch);
    IRB_REQUIRE(B <= 65535, "%s: B too large", who);
    IRB_REQUIRE((size_t)H * W < (size_t)1 << 30, "%s: image too large", who);
    return 0;
}

extern "C" int irb_bg_build(const float* guide, float guide_scale, int B, int H, int W, double sigma_luma, double sigma_chroma,
                            double sigma_spatial, void* ws, size_t ws_bytes, int* nverts_out, void* stream) {
    if (int rc = bg_check(ws, B, H, W, 1, "bg_build")) return rc;
    IRB_REQUIRE(guide != nullptr, "bg_build: null guide");
    IRB_REQUIRE(sigma_luma > 0 && sigma_chroma > 0 && sigma_spatial > 0, "bg_build: sigmas must be positive");
    // the layout's prefix (everything the build touches) does not depend on ch; ws_bytes is checked for ch=1
    const BgLayout L = make_layout(B, H, W, 1);
    if (ws_bytes < L.total) {
        set_error("bg_build: workspace too small (%zu < %zu)", ws_bytes, L.total);
        return IRB_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tot = L.P * (size_t)B;
    cudaMemsetAsync(at<int>(ws, L.oor), 0, sizeof(int) * (size_t)B, st);
    bg_keys_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(guide, guide_scale, ws, L, sigma_luma, sigma_chroma, sigma_spatial);
    // hash < 255^5 < 2^40 whenever every coordinate is < 255 (the reference's own requirement); otherwise
    // sort all 64 bits so that aliasing hashes are still ordered exactly like np.unique orders them
    const double maxc = fmax(fmax((double)W / sigma_spatial, (double)H / sigma_spatial), fmax(256.0 / sigma_luma, 384.0 / sigma_chroma));
    const int npass = maxc < 254.0 ? 5 : 7;  // odd: the sorted result must land in keysB / payB
    cudaMemsetAsync(at<unsigned>(ws, L.dtot), 0, 4 * 256 * (size_t)8 * B, st);
    {
        const dim3 sgrid(SORT_WPI / SORT_CTA_WARPS, (unsigned)B);
        for (int pass = 0; pass < npass; ++pass) {
            bg_sort_hist_kernel<<<sgrid, SORT_CTA_WARPS * 32, 0, st>>>(ws, L, pass);
            bg_sort_scan_kernel<<<dim3(32, (unsigned)B), 256, 0, st>>>(ws, L, pass);
            bg_sort_scatter_kernel<<<sgrid, SORT_CTA_WARPS * 32, 0, st>>>(ws, L, pass);
        }
        bg_unique_count_kernel<<<sgrid, SORT_CTA_WARPS * 32, 0, st>>>(ws, L);
        bg_unique_rank_kernel<<<sgrid, SORT_CTA_WARPS * 32, 0, st>>>(ws, L);
    }
    const dim3 grid((unsigned)((L.P + 255) / 256), (unsigned)B);
    bg_nbr_kernel<<<grid, 256, 0, st>>>(ws, L);
    for (int i = 0; i < 10; ++i) bg_bisto_iter_kernel<<<grid, 256, 0, st>>>(ws, L, i & 1);
    bg_bisto_final_kernel<<<grid, 256, 0, st>>>(ws, L, 0);  // after 10 iterations n is back in n0
    if (int rc = check_launch("bg_build", 1 + 3 * npass + 2 + 1 + 11)) return rc;
    if (nverts_out) {
        cudaError_t e = cudaMemcpyAsync(nverts_out, at<int>(ws, L.hdr), sizeof(int) * B, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("bg_build: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    return 0;
}

extern "C" int irb_bg_vertex_index(const void* ws, int32_t* idx_out, int B, int H, int W, void* stream) {
    if (int rc = bg_check(ws, B, H, W, 1, "bg_vertex_index")) return rc;
    IRB_REQUIRE(idx_out != nullptr, "bg_vertex_index: null output");
    const BgLayout L = make_layout(B, H, W, 1);
    const cudaError_t e = cudaMemcpyAsync(idx_out, at<int>(ws, L.idx), 4 * L.P * (size_t)B, cudaMemcpyDeviceToDevice,
                                          (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("bg_vertex_index: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

extern "C" int irb_bg_solve_fwd(void* ws, const float* pred, const float* conf, int B, int ch, int H, int W, double lam,
                                double A_diag_min, double cg_tol, int cg_maxiter, float* out, double* yhat, void* stream) {
    if (int rc = bg_check(ws, B, H, W, ch, "bg_solve_fwd")) return rc;
    IRB_REQUIRE(pred && conf && out && yhat, "bg_solve_fwd: null pointer");
    IRB_REQUIRE(cg_maxiter >= 0 && cg_maxiter <= BG_MAX_ITERS, "bg_solve_fwd: cg_maxiter must be in [0,%d]", BG_MAX_ITERS);
    const BgLayout L = make_layout(B, H, W, ch);
    cudaStream_t st = (cudaStream_t)stream;
    SolveArgs a{pred, conf, pred, yhat, lam, A_diag_min, cg_tol, 0, 0};
    if (int rc = run_pcg(ws, L, a, cg_maxiter, st)) return rc;
    const dim3 grid((unsigned)((L.P + 255) / 256), (unsigned)B);
    bg_slice_kernel<<<grid, 256, 0, st>>>(ws, L, yhat, out);
    return check_launch("bg_solve_fwd");
}

extern "C" int irb_bg_solve_bwd(void* ws, const float* pred, const float* conf, const float* g_out, const double* yhat,
                                int B, int ch, int H, int W, double lam, double A_diag_min, double cg_tol, int cg_maxiter,
                                float* g_pred, float* g_conf, void* stream) {
    if (int rc = bg_check(ws, B, H, W, ch, "bg_solve_bwd")) return rc;
    IRB_REQUIRE(pred && conf && g_out && yhat, "bg_solve_bwd: null pointer");
    IRB_REQUIRE(cg_maxiter >= 0 && cg_maxiter <= BG_MAX_ITERS, "bg_solve_bwd: cg_maxiter must be in [0,%d]", BG_MAX_ITERS);
    const BgLayout L = make_layout(B, H, W, ch);
    cudaStream_t st = (cudaStream_t)stream;
    double* gh = at<double>(ws, L.sol);
    SolveArgs a{pred, conf, g_out, gh, lam, A_diag_min, cg_tol, 1, 0};
    if (int rc = run_pcg(ws, L, a, cg_maxiter, st)) return rc;
    const dim3 grid((unsigned)((L.P + 255) / 256), (unsigned)B);
    bg_grad_kernel<<<grid, 256, 0, st>>>(ws, L, gh, yhat, pred, conf, g_pred, g_conf);
    return check_launch("bg_solve_bwd");
}

extern "C" int irb_bg_apply(const void* ws, int op, const double* x, double* out, int B, int ch, int H, int W, void* stream) {
    if (int rc = bg_check(ws, B, H, W, ch, "bg_apply")) return rc;
    IRB_REQUIRE(x && out && x != out, "bg_apply: null or aliased pointer");
    IRB_REQUIRE(op >= 0 && op <= 2, "bg_apply: op must be 0 (splat), 1 (slice) or 2 (blur)");
    const BgLayout L = make_layout(B, H, W, ch);
    const dim3 grid((unsigned)((L.P + 255) / 256), (unsigned)B);
    bg_op_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ws, L, op, x, out);
    return check_launch("bg_apply");
}

extern "C" int irb_bg_bistochastization(const void* ws, double* n, double* m, int B, int H, int W, void* stream) {
    if (int rc = bg_check(ws, B, H, W, 1, "bg_bistochastization")) return rc;
    IRB_REQUIRE(n && m, "bg_bistochastization: null pointer");
    const BgLayout L = make_layout(B, H, W, 1);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(n, at<double>(ws, L.n0), 8 * L.P * (size_t)B, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(m, at<double>(ws, L.m), 8 * L.P * (size_t)B, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
        set_error("bg_bistochastization: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

extern "C" int irb_bg_stats(const void* ws, int* nverts, unsigned long long* nlinks, int B, int H, int W, void* stream) {
    if (int rc = bg_check(ws, B, H, W, 1, "bg_stats")) return rc;
    IRB_REQUIRE(nverts && nlinks, "bg_stats: null pointer");
    const BgLayout L = make_layout(B, H, W, 1);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(nlinks, 0, 8 * (size_t)B, st);
    const dim3 grid((unsigned)((L.P + 255) / 256), (unsigned)B);
    bg_links_kernel<<<grid, 256, 0, st>>>(ws, L, nlinks);
    const cudaError_t e = cudaMemcpyAsync(nverts, at<int>(ws, L.hdr), 4 * (size_t)B, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
        set_error("bg_stats: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return check_launch("bg_stats");
}

// debugging / test aid: the per-(image, channel, iteration) PCG scalars {rho, <p,q>, <r,r>, -} of the LAST solve on this
// workspace, device fp64 [B][BG_MAX_CH = 4][BG_MAX_ITERS + 2 = 258][4]; slot 257 holds {<b,b>, -, -, -}
extern "C" int irb_bg_debug_scalars(const void* ws, double* out, int B, int ch, int H, int W, void* stream) {
    if (int rc = bg_check(ws, B, H, W, ch, "bg_debug_scalars")) return rc;
    IRB_REQUIRE(out != nullptr, "bg_debug_scalars: null output");
    const BgLayout L = make_layout(B, H, W, ch);
    const cudaError_t e = cudaMemcpyAsync(out, at<double>(ws, L.scal), 8 * (size_t)B * BG_MAX_CH * (BG_MAX_ITERS + 2) * 4,
                                          cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("bg_debug_scalars: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

extern "C" int irb_bg_range_flags(const void* ws, int* flags, int B, int H, int W, void* stream) {
    if (int rc = bg_check(ws, B, H, W, 1, "bg_range_flags")) return rc;
    IRB_REQUIRE(flags != nullptr, "bg_range_flags: null output");
    const BgLayout L = make_layout(B, H, W, 1);
    const cudaError_t e = cudaMemcpyAsync(flags, at<int>(ws, L.oor), 4 * (size_t)B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("bg_range_flags: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}
