// irb_render_env_tma.cuh -- TMA-staged render_env forward / backward (models.renderingLayer.forwardEnv on a MATERIALISED
// env-map, models.py:466-520).  Device only.
//
// Why: the thread-per-pixel kernels read a pixel's 128 directions as 16-byte vectors at a 512-byte stride between lanes --
// every warp load touches 32 different 128-byte lines and half of each 32-byte sector; ncu r02: backward 62 % long-scoreboard
// stalls and 1.29x the algorithmic DRAM traffic.  Here the env-map (and in the backward its gradient) moves through shared
// memory in tiles of 128 pixels x DC directions x 3 channels (DC = 8 shipped, 16 kept as an A/B variant):
//   * tensor map over env viewed as [B*3, P, D] fp32, box {DC directions, 128 pixels, 1}, swizzle span = the row (32 / 64
//     bytes) -- every row of the box is one full run of a pixel's directions, the hardware XOR-swizzle makes the per-lane
//     16-byte reads (lane = pixel) bank-conflict free;
//   * a 3-stage ring with full mbarriers; the warp that is through with a stage last refills it (no producer thread);
//   * backward: the thread that read E[c][d0..d0+3] writes dL/dE into the same slots, and the stage goes back to HBM as
//     ONE tensor-map store per channel (cp.async.bulk.tensor ... shared -> global, SASS UTMASTG) before it is refilled;
//   * regular pixels: packed f32x2 local-frame BRDF, two directions per instruction (RenderEnvPixelPk below).
// Per-pixel math: RenderEnvPixel of irb_driver_env.cuh (the same code the plain kernels and the host check run).
// Measurements of every step: profiles/r02/ab_render_env.txt.
#pragma once

namespace irb {

constexpr int RE_PT = 128;                       // pixels per tile = threads per CTA
// directions per chunk (16 -> 64-byte rows, SWIZZLE_64B; 8 -> 32-byte rows, SWIZZLE_32B), ring depth and resident CTAs per
// SM, per direction of the pass.  Shared memory per thread = stages x DC x 12 bytes, so DC sets how many warps fit beside
// the ring: the forward (80 registers) is issue-bound and wants 24 warps -> DC 8; see profiles/r02/ab_render_env.txt
#ifndef IRB_RE_FWD_DC
#define IRB_RE_FWD_DC 8
#endif
#ifndef IRB_RE_BWD_DC
#define IRB_RE_BWD_DC 8
#endif
#ifndef IRB_RE_FWD_MINB
#define IRB_RE_FWD_MINB 5
#endif
#ifndef IRB_RE_BWD_MINB
#define IRB_RE_BWD_MINB 3
#endif
#ifndef IRB_RE_STAGES
#define IRB_RE_STAGES 3
#endif
constexpr int RE_S = IRB_RE_STAGES;              // stages
template <bool BWD>
struct ReCfg {
    static constexpr int DC = BWD ? IRB_RE_BWD_DC : IRB_RE_FWD_DC;
    static constexpr int MINB = BWD ? IRB_RE_BWD_MINB : IRB_RE_FWD_MINB;
    static constexpr int CH_BYTES = RE_PT * DC * 4;   // one channel of a chunk
    static constexpr int STAGE_BYTES = 3 * CH_BYTES;
    static_assert(DC == 8 || DC == 16, "rows of 32 or 64 bytes");
};

struct EnvTileMaps {
    CUtensorMap env, genv;
};

// env viewed as [nb3 = B*3, P, D] fp32; box {dc, RE_PT, 1}; the swizzle span equals the row (dc * 4 bytes)
inline bool make_env_tile_map(CUtensorMap* m, const float* base, size_t P, unsigned D, size_t nb3, unsigned dc) {
    irb_encode_tiled_fn fn = tile_map_encoder();
    if (!fn || D % dc != 0 || ((uintptr_t)base & 15u) || P >= 0x7fffffffull || nb3 >= 0x7fffffffull) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)P, (cuuint64_t)nb3};
    const cuuint64_t strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)D * 4 * P};
    const cuuint32_t box[3] = {(cuuint32_t)dc, (cuuint32_t)RE_PT, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, dc == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
              CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// byte offset of the 16-byte block `j` (4 directions) of pixel-row `row` inside one channel of a stage.  The swizzle
// pattern repeats every 8 rows (SWIZZLE_64B: 512 bytes, address bits [5:4] ^= bits [8:7]; SWIZZLE_32B: 256 bytes, bit 4 ^=
// bit 7), stage bases are 1024-byte aligned.  Either way lane = row reads are bank-conflict free.
template <int DC>
__device__ __forceinline__ int re_off(int row, int j) {
    const int x = DC == 16 ? ((row >> 1) & 3) : ((row >> 2) & 1);
    return row * (DC * 4) + ((j ^ x) << 4);
}

// per-thread cache of the two descriptors (same idea as cached_tile_map)
inline bool cached_env_tile_map(CUtensorMap* m, const float* base, size_t P, unsigned D, size_t nb3, unsigned dc) {
    struct Entry {
        const float* base;
        size_t P, nb3;
        unsigned D, dc;
        bool valid;
        CUtensorMap map;
    };
    static thread_local Entry tab[16];
    const size_t h = ((uintptr_t)base >> 4) * 0x9e3779b97f4a7c15ull ^ P * 31 ^ nb3 * 131 ^ D ^ dc * 7;
    Entry& e = tab[(h >> 20) % 16];
    if (e.valid && e.base == base && e.P == P && e.nb3 == nb3 && e.D == D && e.dc == dc) {
        memcpy(m, &e.map, sizeof(CUtensorMap));
        return true;
    }
    if (!make_env_tile_map(m, base, P, D, nb3, dc)) return false;
    e.base = base; e.P = P; e.nb3 = nb3; e.D = D; e.dc = dc; e.valid = true;
    memcpy(&e.map, m, sizeof(CUtensorMap));
    return true;
}

// Packed (two directions per instruction) per-pixel state for regular pixels: the same local-frame formulas as
// RenderEnvPixel::quad_fwd / quad_bwd, evaluated with dir_fwd_local2 / dir_bwd_local2 (irb_pixel.cuh) on the pairs
// (d, d+1) -- the float4 a thread reads from the stage already holds them in adjacent registers.  Irregular pixels
// (normal close to zero ...) take the scalar world-space path of the base class.  Even / odd directions are summed
// separately and added at the end (fp32 rounding only; parity bar: tests/test_render_gpu.py).
#ifndef IRB_RE_PK_FWD
#define IRB_RE_PK_FWD 1
#endif
#ifndef IRB_RE_PK_BWD
#define IRB_RE_PK_BWD 1
#endif
struct RenderEnvPixelPk : RenderEnvPixel<float> {
    pk2 Dc2[3], Sc2[3], gAlb2[3];
    PixelGradL2 gl2;

    __device__ __forceinline__ void begin_pk(const RenderEnvArgs& s, size_t b, size_t pix, size_t plane, bool bwd) {
        begin(s, b, pix, plane, bwd);
#pragma unroll
        for (int c = 0; c < 3; ++c) Dc2[c] = Sc2[c] = gAlb2[c] = pk(0.f, 0.f);
        if (bwd) pixel_gradl2_zero(gl2);
    }
    // e[c] = env[c][d0 .. d0+3]
    __device__ __forceinline__ void quad_fwd_pk(const EnvTab& tab, int d0, const float4 (&e)[3]) {
        if (!px.regular) {
            const float E[3][4] = {{e[0].x, e[0].y, e[0].z, e[0].w}, {e[1].x, e[1].y, e[1].z, e[1].w}, {e[2].x, e[2].y, e[2].z, e[2].w}};
            quad_fwd(tab, d0, E);
            return;
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int d = d0 + 2 * p;
            const pk2 wx = pk(tab.lx[d], tab.lx[d + 1]), wy = pk(tab.ly[d], tab.ly[d + 1]), wz = pk(tab.lz[d], tab.lz[d + 1]);
            DirL2 dl;
            dir_fwd_local2(dl, px, pl, wx, wy, wz);
            const pk2 w = mul2(pk(tab.w[d], tab.w[d + 1]), wz);
            const pk2 ws = mul2(w, dl.spec);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const pk2 E2 = p == 0 ? pk(e[c].x, e[c].y) : pk(e[c].z, e[c].w);
                Dc2[c] = fma2(w, E2, Dc2[c]);
                Sc2[c] = fma2(ws, E2, Sc2[c]);
            }
        }
    }
    __device__ __forceinline__ void end_fwd_pk(const RenderEnvArgs& s, size_t b, size_t pix, size_t plane) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Dc[c] += hsum2(Dc2[c]);
            Sc[c] += hsum2(Sc2[c]);
        }
        end_fwd(s, b, pix, plane);
    }
    // ge[c] = dL/d env[c][d0 .. d0+3]
    __device__ __forceinline__ void quad_bwd_pk(const EnvTab& tab, int d0, const float4 (&e)[3], float4 (&ge)[3]) {
        if (!px.regular) {
            const float E[3][4] = {{e[0].x, e[0].y, e[0].z, e[0].w}, {e[1].x, e[1].y, e[1].z, e[1].w}, {e[2].x, e[2].y, e[2].z, e[2].w}};
            float GE[3][4];
            quad_bwd(tab, d0, E, GE);
#pragma unroll
            for (int c = 0; c < 3; ++c) ge[c] = make_float4(GE[c][0], GE[c][1], GE[c][2], GE[c][3]);
            return;
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int d = d0 + 2 * p;
            const pk2 wx = pk(tab.lx[d], tab.lx[d + 1]), wy = pk(tab.ly[d], tab.ly[d + 1]), wz = pk(tab.lz[d], tab.lz[d + 1]);
            DirL2 dl;
            dir_fwd_local2(dl, px, pl, wx, wy, wz);
            const pk2 w = mul2(pk(tab.w[d], tab.w[d + 1]), wz);      // n.l = w_d.z: constant, no gradient through it
            pk2 eS = pk(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const pk2 E2 = p == 0 ? pk(e[c].x, e[c].y) : pk(e[c].z, e[c].w);
                eS = fma2(pkb(gS[c]), E2, eS);
                gAlb2[c] = fma2(w, E2, gAlb2[c]);
                const pk2 g2 = mul2(w, fma2(pkb(gS[c]), dl.spec, pkb(gDa[c])));
                if (p == 0) upk(g2, ge[c].x, ge[c].y);
                else upk(g2, ge[c].z, ge[c].w);
            }
            dir_bwd_local2(gl2, px, pl, dl, wx, wy, wz, mul2(w, eS));
        }
    }
    __device__ __forceinline__ void end_bwd_pk(const RenderEnvArgs& s, size_t b, size_t pix, size_t plane) {
        if (px.regular) {
            pixel_gradl2_fold(gl, gl2);
#pragma unroll
            for (int c = 0; c < 3; ++c) gAlb[c] += hsum2(gAlb2[c]);
        }
        end_bwd(s, b, pix, plane);
    }
};

template <bool BWD>
__global__ void __launch_bounds__(RE_PT, ReCfg<BWD>::MINB)
    render_env_tma_kernel(const RenderEnvArgs s, const __grid_constant__ EnvTab tab, const __grid_constant__ EnvTileMaps maps,
                          const unsigned total_tiles, const unsigned tiles_per_img) {
    constexpr int RE_DC = ReCfg<BWD>::DC, RE_CH_BYTES = ReCfg<BWD>::CH_BYTES, RE_STAGE_BYTES = ReCfg<BWD>::STAGE_BYTES;
    // no static shared memory in this kernel: the dynamic window starts at the CTA's (1 KB-granular) allocation, the stages
    // come first so that their swizzle atoms are aligned, the barriers follow.  Checked, not assumed.
    extern __shared__ __align__(1024) unsigned char re_smem[];
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(re_smem + RE_S * RE_STAGE_BYTES);
    unsigned* const done_cnt = reinterpret_cast<unsigned*>(full_bar + RE_S);   // warps that are through with the stage
    if (threadIdx.x == 0) {
        if (smem_u32(re_smem) & 1023u) __trap();
#pragma unroll
        for (int i = 0; i < RE_S; ++i) {
            mbar_init(&full_bar[i], 1);
            done_cnt[i] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t plane = (size_t)s.R * s.C;
    const int D = tab.D, nchunk = D / RE_DC;
    const bool store = BWD && s.g_env != nullptr;
    const unsigned my_tiles = blockIdx.x < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const unsigned G = my_tiles * (unsigned)nchunk;   // chunks this CTA walks
    auto chunk_coords = [&](unsigned g, int& d0, int& pix0, int& b) {
        const unsigned t = blockIdx.x + (g / (unsigned)nchunk) * gridDim.x;
        d0 = (int)(g % (unsigned)nchunk) * RE_DC;
        b = (int)(t / tiles_per_img);
        pix0 = (int)(t % tiles_per_img) * RE_PT;
    };
    auto issue = [&](unsigned g) {   // elected thread: fill stage g % RE_S with chunk g
        int d0, pix0, b;
        chunk_coords(g, d0, pix0, b);
        unsigned char* st = re_smem + (size_t)(g % RE_S) * RE_STAGE_BYTES;
        uint64_t* bar = &full_bar[g % RE_S];
        mbar_expect_tx(bar, (uint32_t)RE_STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_load_3d(st + c * RE_CH_BYTES, &maps.env, d0, pix0, b * 3 + c, bar);
    };
    auto flush = [&](unsigned g) {   // elected thread: stage g % RE_S holds dL/dE of chunk g -> global
        int d0, pix0, b;
        chunk_coords(g, d0, pix0, b);
        const unsigned char* st = re_smem + (size_t)(g % RE_S) * RE_STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_store_3d(&maps.genv, st + c * RE_CH_BYTES, d0, pix0, b * 3 + c);
        tma_store_commit();
    };
    if (threadIdx.x == 0)
        for (unsigned g = 0; g < (unsigned)RE_S && g < G; ++g) issue(g);

    constexpr bool PK = BWD ? (IRB_RE_PK_BWD != 0) : (IRB_RE_PK_FWD != 0);
    RenderEnvPixelPk px;
    const int row = threadIdx.x;
    int d0 = 0, pix0 = 0, b = 0;   // chunk g's coordinates, advanced incrementally (one division pair per tile)
    unsigned tile = blockIdx.x;
    if (G > 0) {
        b = (int)(tile / tiles_per_img);
        pix0 = (int)(tile % tiles_per_img) * RE_PT;
    }
    for (unsigned g = 0; g < G; ++g) {
        const int stage = (int)(g % RE_S);
        const bool active = (size_t)pix0 + row < plane;
        const size_t pix = active ? (size_t)pix0 + row : (size_t)pix0;   // idle lanes shade the tile's first pixel, nothing stored
        if (d0 == 0) px.begin_pk(s, (size_t)b, pix, plane, BWD);
        mbar_wait(&full_bar[stage], (g / RE_S) & 1);
        unsigned char* st = re_smem + (size_t)stage * RE_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < RE_DC / 4; ++j) {
            const int off = re_off<RE_DC>(row, j);
            const float4 e0 = *reinterpret_cast<const float4*>(st + off);
            const float4 e1 = *reinterpret_cast<const float4*>(st + RE_CH_BYTES + off);
            const float4 e2 = *reinterpret_cast<const float4*>(st + 2 * RE_CH_BYTES + off);
            if constexpr (PK) {
                const float4 e[3] = {e0, e1, e2};
                if (BWD) {
                    float4 ge[3];
                    px.quad_bwd_pk(tab, d0 + 4 * j, e, ge);
                    if (store) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(st + c * RE_CH_BYTES + off) = ge[c];
                    }
                } else {
                    px.quad_fwd_pk(tab, d0 + 4 * j, e);
                }
                continue;
            }
            const float E[3][4] = {{e0.x, e0.y, e0.z, e0.w}, {e1.x, e1.y, e1.z, e1.w}, {e2.x, e2.y, e2.z, e2.w}};
            if (BWD) {
                float GE[3][4];
                px.quad_bwd(tab, d0 + 4 * j, E, GE);
                if (store) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        *reinterpret_cast<float4*>(st + c * RE_CH_BYTES + off) = make_float4(GE[c][0], GE[c][1], GE[c][2], GE[c][3]);
                }
            } else {
                px.quad_fwd(tab, d0 + 4 * j, E);
            }
        }
        // The warp that is through with the stage LAST sends its gradient tile home (backward) and refills it with chunk
        // g + RE_S: no designated producer thread, so no warp waits for a particular other warp (with thread 0 as the
        // producer warp 0 trailed every chunk and the others queued behind its refills).
        if (store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the bulk store
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            unsigned old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&done_cnt[stage])) : "memory");
            if (old == (unsigned)(RE_PT / 32 - 1)) {
                done_cnt[stage] = 0u;   // untouched until the refill below has landed
                if (store) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    flush(g);
                    if (g + RE_S < G) tma_store_wait_read();   // the bulk store has finished READING the stage
                }
                if (g + RE_S < G) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(g + RE_S);
                }
            }
        }
        if (d0 + RE_DC >= D && active) {
            if constexpr (!PK) {
                if (BWD) px.end_bwd(s, (size_t)b, pix, plane);
                else px.end_fwd(s, (size_t)b, pix, plane);
            } else if constexpr (BWD) {
                px.end_bwd_pk(s, (size_t)b, pix, plane);
            } else {
                px.end_fwd_pk(s, (size_t)b, pix, plane);
            }
        }
        d0 += RE_DC;
        if (d0 >= D) {
            d0 = 0;
            tile += gridDim.x;
            b = (int)(tile / tiles_per_img);
            pix0 = (int)(tile % tiles_per_img) * RE_PT;
        }
    }
    if (store) tma_store_wait_all();   // threads that never stored return at once
}

}  // namespace irb
