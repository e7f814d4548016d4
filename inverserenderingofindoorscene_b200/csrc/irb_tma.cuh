// irb_tma.cuh -- bulk-async-copy (TMA) and mbarrier primitives shared by the persistent, smem-staged kernels.
#pragma once
#include <stdint.h>

#include <cuda.h>   // CUtensorMap and its enums only; the encoder is fetched with cudaGetDriverEntryPoint (no -lcuda)

namespace irb {

// Tensor maps of the planes a tile stages, in stage order.  on = 0: descriptors unavailable, kernels fall back to one
// cp.async.bulk per plane.
constexpr int kMaxTileMaps = 10;
struct TileMaps {
    CUtensorMap m[kMaxTileMaps];
    int on;
};

typedef CUresult (*irb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline irb_encode_tiled_fn tile_map_encoder() {
    static irb_encode_tiled_fn fn = []() -> irb_encode_tiled_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (irb_encode_tiled_fn)p;
    }();
    return fn;
}

// fp32 tensor [nb, rows, plane] (contiguous), box = [1, rows, tile] -> false when it cannot be described
inline bool make_tile_map(CUtensorMap* m, const float* base, size_t plane, unsigned rows, unsigned nb, unsigned tile) {
    irb_encode_tiled_fn fn = tile_map_encoder();
    if (!fn || plane < tile || rows == 0 || rows > 256 || tile > 256 || plane % 4 != 0 || ((uintptr_t)base & 15u)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)plane, rows, nb};
    const cuuint64_t strides[2] = {(cuuint64_t)plane * 4, (cuuint64_t)plane * 4 * rows};
    const cuuint32_t box[3] = {tile, rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

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
// tiled (tensor-map) form: one instruction moves a [rows x PT] box of a [B, rows, plane] tensor; elements past the end
// of the plane are zero-filled and still counted in the transaction bytes (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
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
