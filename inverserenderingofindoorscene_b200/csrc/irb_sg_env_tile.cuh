// irb_sg_env_tile.cuh -- sg_to_env forward (models.output2env.fromSGtoIm, models.py:388-407) with the env tile assembled in
// shared memory and stored with tensor-map stores.  Device only.
//
// Why (ncu r02, sg_env_group_kernel<0,2> at 4x240x320, K=12: 351 us, XU 35 %, FMA 33 %, short-scoreboard 26 %): the
// lobe-split kernel gives a pixel to TPP lanes and adds their partial sums with 3 shuffles per direction, and each lane
// stores 16 bytes at a 512-byte stride.  The kernel has no BRDF chain -- nothing needs the lobes split -- so here ONE thread
// owns a pixel and ALL its K lobes (K = 6 / 12: 42 / 84 registers as f32x2 pairs): no shuffles, 12 independent MUFU.EX2 per
// direction in flight, and the [3 ch][128 px][16 dir] tile leaves through a 2-stage shared-memory ring as UTMASTG stores
// (same swizzled layout as irb_render_env_tma.cuh).  Bound: the MUFU pipe -- K x D exponentials per pixel = 105 us at
// 16 / clk / SM for this shape -- above the 65 us of the 472 MB the kernel has to write.
#pragma once

namespace irb {

#ifndef IRB_SGT_DC
#define IRB_SGT_DC 16
#endif
#ifndef IRB_SGT_MINB
#define IRB_SGT_MINB 4
#endif
#ifndef IRB_SGT_STAGES
#define IRB_SGT_STAGES 2
#endif

template <int N> __device__ __forceinline__ void tma_store_wait_read_n() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int KC>
__global__ void __launch_bounds__(RE_PT, IRB_SGT_MINB)
    sg_env_tile_fwd_kernel(const SgEnvArgs s, const __grid_constant__ EnvTab tab, const __grid_constant__ CUtensorMap emap,
                           const unsigned total_tiles, const unsigned tiles_per_img) {
    constexpr int DC = IRB_SGT_DC, CH_BYTES = RE_PT * DC * 4, STAGE_BYTES = 3 * CH_BYTES, S = IRB_SGT_STAGES, NP2 = KC / 2;
    constexpr int NWARP = RE_PT / 32;
    static_assert(DC == 8 || DC == 16, "rows of 32 or 64 bytes");
    extern __shared__ __align__(1024) unsigned char sgt_smem[];   // stages first (swizzle atoms aligned), then the barriers
    uint64_t* const free_bar = reinterpret_cast<uint64_t*>(sgt_smem + S * STAGE_BYTES);   // the stage's store has read it out
    unsigned* const done_cnt = reinterpret_cast<unsigned*>(free_bar + S);                 // warps that finished writing the stage
    if (threadIdx.x == 0) {
        if (smem_u32(sgt_smem) & 1023u) __trap();
#pragma unroll
        for (int i = 0; i < S; ++i) {
            mbar_init(&free_bar[i], 1);
            done_cnt[i] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t plane = (size_t)s.R * s.C;
    const int D = tab.D;
    const int row = threadIdx.x;
    unsigned g = 0;   // chunks this CTA has produced so far
    // whole tiles, round-robin.  (A chunk-granular balanced split -- every CTA U/G +- 1 chunks, shared tiles loading their
    // lobes twice -- was measured: 229 vs 212 us; the few CTAs of the last round run alone on their SM and finish fast.)
    for (unsigned tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_img), pix0 = (int)(tile % tiles_per_img) * RE_PT;
        const bool active = (size_t)pix0 + row < plane;
        const size_t pix = active ? (size_t)pix0 + row : (size_t)pix0;   // idle lanes redo the tile's first pixel; the store clips
        // ---- the pixel's K lobes.  Two phases on purpose: all 7K loads are issued before the first tan warp so that the
        // tile pays ONE global-memory latency (ncu r02: with load/warp interleaved per lobe, 31 % of the kernel's stall samples
        // sat on the first use after each load)
        LobesPk<KC> Lp;
        pk2 W2[3][NP2];
        {
            float ax[KC][3], lm[KC], W[KC][3];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const size_t oa = (((size_t)b * KC + k) * 3) * plane + pix;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    ax[k][c] = __ldg(s.axis + oa + c * plane);
                    W[k][c] = __ldg(s.weight + ((size_t)b * (3 * KC) + 3 * k + c) * plane + pix);
                }
                lm[k] = __ldg(s.lamb + ((size_t)b * KC + k) * plane + pix);
            }
            Lobes<float, KC> L;
#pragma unroll
            for (int k = 0; k < KC; ++k) {   // same arithmetic as load_lobes / load_weight (irb_driver.cuh)
                float lam = lm[k];
                if (s.warp) lam = sg_warp<float>(lam);
                lam *= Consts<float>::LOG2E;
                L.lam[k] = lam;
                L.bx[k] = lam * ax[k][0];
                L.by[k] = lam * ax[k][1];
                L.bz[k] = lam * ax[k][2];
                if (s.warp) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) W[k][c] = sg_warp<float>(W[k][c]);
                }
            }
            pack_lobes<KC>(Lp, L);
            if (active && (s.lamb_w || s.weight_w)) {
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    if (s.lamb_w) s.lamb_w[((size_t)b * KC + k) * plane + pix] = L.lam[k] * Consts<float>::LN2;
                    if (s.weight_w)
#pragma unroll
                        for (int c = 0; c < 3; ++c) s.weight_w[((size_t)b * (3 * KC) + 3 * k + c) * plane + pix] = W[k][c];
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int i = 0; i < NP2; ++i) W2[c][i] = pk(W[2 * i][c], W[2 * i + 1][c]);
        }
        for (int d0 = 0; d0 < D; d0 += DC, ++g) {
            const int stage = (int)(g % S);
            if (g >= (unsigned)S) mbar_wait(&free_bar[stage], ((g / S) - 1) & 1);   // chunk g - S has left the stage
            unsigned char* st = sgt_smem + (size_t)stage * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < DC / 4; ++j) {
                float E[3][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int d = d0 + 4 * j + u;
                    const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                    pk2 arg[NP2], e[NP2];
                    lobe_exps_pk<KC>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg, e);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        pk2 acc = mul2(W2[c][0], e[0]);
#pragma unroll
                        for (int i = 1; i < NP2; ++i) acc = fma2(W2[c][i], e[i], acc);
                        E[c][u] = hsum2(acc);
                    }
                }
                const int off = re_off<DC>(row, j);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    *reinterpret_cast<float4*>(st + c * CH_BYTES + off) = make_float4(E[c][0], E[c][1], E[c][2], E[c][3]);
            }
            // The warp that finishes the stage last sends it home: no warp is the designated flusher, so nobody waits for
            // a particular warp (a fixed flusher made the CTA run in lock-step: 21 try-wait spins per chunk in ncu).
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the bulk store
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {
                unsigned old;
                asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&done_cnt[stage])) : "memory");
                if (old == (unsigned)(NWARP - 1)) {
                    done_cnt[stage] = 0u;   // nobody touches it again before free_bar flips
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 3; ++c) tma_store_3d(&emap, st + c * CH_BYTES, d0, pix0, b * 3 + c);
                    tma_store_commit();
                    tma_store_wait_read();   // this thread's stores have read the stage
                    mbar_arrive(&free_bar[stage]);
                }
            }
        }
    }
    tma_store_wait_all();   // threads that never stored return at once
}

// ---- reconstruction statistics / loss against a ground-truth env-map, thread per pixel ----------------------------------
// MODE 2: per-image regression sums  sum(E G s^2), sum(E^2 s^2), sum(s)     (models.LSregress, models.py:7-21)
// MODE 3: numerator  sum((log(c E + off) - log(G + off))^2 s)               (wrapperBRDFLight.py:183-187)
// Same arithmetic per (pixel, direction, channel) as sg_env_group_kernel<2|3,TPP>; here the predicted radiance E is formed
// like in sg_env_tile_fwd_kernel (all K lobes in one thread) and the ground truth arrives through a 3-stage TMA ring of
// [3 ch][128 px][8 dir] tiles (coalesced; the group kernel reads 16 bytes per lane at a 512-byte stride).
#ifndef IRB_SGT_STAT_MINB
#define IRB_SGT_STAT_MINB 4
#endif
template <int KC, int MODE>
__global__ void __launch_bounds__(RE_PT, IRB_SGT_STAT_MINB)
    sg_env_tile_stat_kernel(const SgEnvArgs s, const __grid_constant__ EnvTab tab, const __grid_constant__ CUtensorMap gmap,
                            const unsigned total_tiles, const unsigned tiles_per_img) {
    constexpr int DC = 8, CH_BYTES = RE_PT * DC * 4, STAGE_BYTES = 3 * CH_BYTES, S = 3, NP2 = KC / 2, NWARP = RE_PT / 32;
    extern __shared__ __align__(1024) unsigned char sgs_smem[];
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(sgs_smem + S * STAGE_BYTES);
    unsigned* const done_cnt = reinterpret_cast<unsigned*>(full_bar + S);
    if (threadIdx.x == 0) {
        if (smem_u32(sgs_smem) & 1023u) __trap();
#pragma unroll
        for (int i = 0; i < S; ++i) {
            mbar_init(&full_bar[i], 1);
            done_cnt[i] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t plane = (size_t)s.R * s.C;
    const int D = tab.D, nchunk = D / DC;
    const int row = threadIdx.x;
    const unsigned my_tiles = blockIdx.x < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const unsigned G = my_tiles * (unsigned)nchunk;
    auto issue = [&](unsigned g) {   // fill stage g % S with chunk g of the ground truth
        const unsigned t = blockIdx.x + (g / (unsigned)nchunk) * gridDim.x;
        const int d0 = (int)(g % (unsigned)nchunk) * DC, b = (int)(t / tiles_per_img), pix0 = (int)(t % tiles_per_img) * RE_PT;
        unsigned char* st = sgs_smem + (size_t)(g % S) * STAGE_BYTES;
        uint64_t* bar = &full_bar[g % S];
        mbar_expect_tx(bar, (uint32_t)STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_load_3d(st + c * CH_BYTES, &gmap, d0, pix0, b * 3 + c, bar);
    };
    if (threadIdx.x == 0)
        for (unsigned g = 0; g < (unsigned)S && g < G; ++g) issue(g);
    unsigned g = 0;
    for (unsigned tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_img), pix0 = (int)(tile % tiles_per_img) * RE_PT;
        const bool active = (size_t)pix0 + row < plane;
        const size_t pix = active ? (size_t)pix0 + row : (size_t)pix0;
        LobesPk<KC> Lp;
        pk2 W2[3][NP2];
        {
            float ax[KC][3], lm[KC], W[KC][3];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const size_t oa = (((size_t)b * KC + k) * 3) * plane + pix;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    ax[k][c] = __ldg(s.axis + oa + c * plane);
                    W[k][c] = __ldg(s.weight + ((size_t)b * (3 * KC) + 3 * k + c) * plane + pix);
                }
                lm[k] = __ldg(s.lamb + ((size_t)b * KC + k) * plane + pix);
            }
            Lobes<float, KC> L;
#pragma unroll
            for (int k = 0; k < KC; ++k) {   // same arithmetic as load_lobes / load_weight (irb_driver.cuh)
                float lam = lm[k];
                if (s.warp) lam = sg_warp<float>(lam);
                lam *= Consts<float>::LOG2E;
                L.lam[k] = lam;
                L.bx[k] = lam * ax[k][0];
                L.by[k] = lam * ax[k][1];
                L.bz[k] = lam * ax[k][2];
                if (s.warp) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) W[k][c] = sg_warp<float>(W[k][c]);
                }
            }
            pack_lobes<KC>(Lp, L);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int i = 0; i < NP2; ++i) W2[c][i] = pk(W[2 * i][c], W[2 * i + 1][c]);
        }
        const float sg = active ? s.seg[(size_t)b * plane + pix] : 0.f;
        const float coef = (MODE == 3) ? reconst_coef(s.stats + (size_t)b * 4) : 1.f;
        float acc0 = 0.f, acc1 = 0.f;
        for (int d0 = 0; d0 < D; d0 += DC, ++g) {
            const int stage = (int)(g % S);
            mbar_wait(&full_bar[stage], (g / S) & 1);
            const unsigned char* st = sgs_smem + (size_t)stage * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < DC / 4; ++j) {
                const int off = re_off<DC>(row, j);
                float4 gq[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) gq[c] = *reinterpret_cast<const float4*>(st + c * CH_BYTES + off);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int d = d0 + 4 * j + u;
                    const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                    pk2 arg[NP2], e[NP2];
                    lobe_exps_pk<KC>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg, e);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        pk2 acc = mul2(W2[c][0], e[0]);
#pragma unroll
                        for (int i = 1; i < NP2; ++i) acc = fma2(W2[c][i], e[i], acc);
                        const float E = hsum2(acc);
                        const float Gv = u == 0 ? gq[c].x : (u == 1 ? gq[c].y : (u == 2 ? gq[c].z : gq[c].w));
                        if (MODE == 2) {
                            acc0 = fmaf(E, Gv, acc0);
                            acc1 = fmaf(E, E, acc1);
                        } else {
                            const float dl = fast_log(coef * E + s.offset) - fast_log(Gv + s.offset);
                            acc0 = fmaf(dl, dl, acc0);
                        }
                    }
                }
            }
            // the warp that is through with the stage last refills it with chunk g + S (irb_render_env_tma.cuh)
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {
                unsigned old;
                asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&done_cnt[stage])) : "memory");
                if (old == (unsigned)(NWARP - 1)) {
                    done_cnt[stage] = 0u;
                    if (g + S < G) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(g + S);
                    }
                }
            }
        }
        // per-pixel weights as in the group kernel: the regression sees (E s)(G s) and (E s)^2, the loss dlog^2 * s;
        // pixels outside the mask contribute exactly zero there too (s = 0)
        double* stt = s.stats + (size_t)b * 4;
        const bool in = sg != 0.f;   // like the group kernel: a masked-out pixel's values are never looked at (NaN-safe)
        if (MODE == 2) {
            sg_block_add(in ? acc0 * sg * sg : 0.f, stt + 0);
            sg_block_add(in ? acc1 * sg * sg : 0.f, stt + 1);
            sg_block_add(sg, stt + 3);
        } else {
            sg_block_add(in ? acc0 * sg : 0.f, stt + 2);
        }
    }
}

}  // namespace irb
