// irb_tma.cuh -- bulk-async-copy (TMA) and mbarrier primitives shared by the persistent, smem-staged kernels.
#pragma once
#include <stdint.h>

namespace irb {

// ---------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (PTX; SASS: UBLKCP, SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && spin > (1u << 26)) __trap();  // a lost copy must fail the launch, never hang the GPU
    }
}


// ---- lobe-split groups: TPP adjacent lanes share one pixel, kLPT lobes per lane -------------------------------
constexpr int kLPT = 6;

template <int TPP> __device__ __forceinline__ float grp_allreduce(float v) {
#pragma unroll
    for (int o = 1; o < TPP; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// p[j] = this lane's partial for the direction owned by group member j; returns sum over members of p[h]
template <int TPP> __device__ __forceinline__ float grp_reduce_scatter(const float (&p)[TPP], int h) {
    if (TPP == 1) return p[0];
    if (TPP == 2) {
        const float send = h ? p[0] : p[1 % TPP];
        const float keep = h ? p[1 % TPP] : p[0];
        return keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    // TPP == 4
    const bool hi = (h & 2) != 0;
    const float a0 = (hi ? p[2 % TPP] : p[0]) + __shfl_xor_sync(0xffffffffu, hi ? p[0] : p[2 % TPP], 2);
    const float a1 = (hi ? p[3 % TPP] : p[1 % TPP]) + __shfl_xor_sync(0xffffffffu, hi ? p[1 % TPP] : p[3 % TPP], 2);
    const bool odd = (h & 1) != 0;
    return (odd ? a1 : a0) + __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
}


}  // namespace irb
