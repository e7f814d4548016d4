// irb_sg_env_group.cuh -- persistent, TMA-staged, lobe-split kernels for SG lobes <-> env-map
// (models.output2env.fromSGtoIm / output2env and their autograd, models.py:371-404).  Device only.
//
// Same structure as the fused render kernels (irb_render_sg_group.cuh): TPP adjacent lanes share a pixel, each
// owns 6 lobes as packed f32x2 pairs; the 7K SG planes of a tile of 128/TPP pixels are staged into shared memory by
// cp.async.bulk one tile ahead.  Forward: every lane evaluates its lobes for all D directions, the three channel
// sums are completed across the group with a butterfly shuffle, and the env-map is written as 16-byte vectors
// (4 directions) straight from registers.  Backward: d/dW, d/dlam, d/daxis accumulate in registers while the
// env-map gradient streams in as 16-byte vectors, loaded one group of 4 directions ahead of its use.
#pragma once

#ifndef IRB_SGENV_FWD_MINB
#define IRB_SGENV_FWD_MINB 4
#endif
#ifndef IRB_SGENV_BWD_MINB
#define IRB_SGENV_BWD_MINB 3
#endif
#ifndef IRB_SGENV_RECONST_MINB
#define IRB_SGENV_RECONST_MINB 2   // loss + gradient sweep: 250 registers, no spills; 573 -> 535 us (profiles/r02/ab_occupancy.txt)
#endif

namespace irb {

__device__ __forceinline__ void issue_sg_tile(float* stage, uint64_t* bar, const SgEnvArgs& s, const TileMaps& maps, size_t tile,
                                              size_t tiles_per_img, int PT) {
    const int lane = threadIdx.x & 31;
    const size_t plane = (size_t)s.R * s.C;
    const size_t b = tile / tiles_per_img, pix0 = (tile % tiles_per_img) * (size_t)PT;
    if (maps.on) {   // one tiled copy per tensor (irb_render_sg_group.cuh)
        if (lane == 0) {
            const int K = s.K;
            mbar_expect_tx(bar, (uint32_t)(7 * K) * (uint32_t)PT * 4u);
            tma_load_3d(stage, &maps.m[0], (int)pix0, 0, (int)b, bar);
            tma_load_3d(stage + 3 * K * PT, &maps.m[1], (int)pix0, 0, (int)b, bar);
            tma_load_3d(stage + 4 * K * PT, &maps.m[2], (int)pix0, 0, (int)b, bar);
        }
        return;
    }
    const uint32_t valid = (uint32_t)((plane - pix0) < (size_t)PT ? (plane - pix0) : (size_t)PT);
    const uint32_t bytes = valid * 4u;
    const int K = s.K, np = 7 * K;
    if (lane == 0) mbar_expect_tx(bar, bytes * (uint32_t)np);
    __syncwarp();
    for (int p = lane; p < np; p += 32) {
        const float* src;
        if (p < 3 * K) src = s.axis + (b * 3 * K + p) * plane;
        else if (p < 4 * K) src = s.lamb + (b * K + (p - 3 * K)) * plane;
        else src = s.weight + (b * 3 * K + (p - 4 * K)) * plane;
        bulk_g2s(stage + (size_t)p * PT, src + pix0, bytes, bar);
    }
}

// MODE 0: forward (env, optional warped lamb / weight), MODE 1: backward from a given env-map gradient,
// MODE 2 / 3 / 4: fused reconstruction loss against a ground-truth env-map WITHOUT materialising the predicted one
//   2: per-sample regression sums  sum(E G s^2), sum(E^2 s^2), sum(s)           (models.LSregress, models.py:7-21)
//   3: numerator  sum((log(c E + off) - log(G + off))^2 s)                        (wrapperBRDFLight.py:183-187)
//   4: its gradient w.r.t. the SG parameters (c is a graph constant, models.py:13)
//   5: 3 and 4 in one pass with a unit upstream gradient (the caller scales the small SG gradients afterwards)
__device__ __forceinline__ void sg_block_add(float v, double* dst) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dst, (double)v);
}
__device__ __forceinline__ float reconst_coef(const double* st) {
    return (float)fmin(fmax(st[0] / fmax(st[1], 1e-5), 0.001), 1000.0);
}
__device__ __forceinline__ float fast_log(float x) {   // natural log via lg2.approx (abs err ~1e-7 for x >= offset > 0)
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.69314718055994530942f;
}

template <int MODE, int TPP>
__global__ void __launch_bounds__(128, (MODE == 1 ? IRB_SGENV_BWD_MINB : (MODE >= 4 ? IRB_SGENV_RECONST_MINB : IRB_SGENV_FWD_MINB))) sg_env_group_kernel(const SgEnvArgs s, const __grid_constant__ EnvTab tab,
                                                              const __grid_constant__ TileMaps maps,
                                                              const unsigned total_tiles, const unsigned tiles_per_img) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2];
    constexpr int PT = 128 / TPP;
    constexpr int NP2 = kLPT / 2;
    const int K = s.K, np = 7 * K, D = tab.D;
    float* const stage0 = reinterpret_cast<float*>(smem_raw);
    float* const stage1 = stage0 + (size_t)np * PT;
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t plane = (size_t)s.R * s.C;
    const int q = threadIdx.x / TPP, h = threadIdx.x % TPP;
    unsigned t = blockIdx.x;
    if (t < total_tiles && threadIdx.x < 32) issue_sg_tile(stage0, &full_bar[0], s, maps, t, tiles_per_img, PT);
    for (unsigned it = 0; t < total_tiles; ++it, t += gridDim.x) {
        const int sidx = it & 1;
        float* const cur = sidx ? stage1 : stage0;
        const unsigned tn = t + gridDim.x;
        if (tn < total_tiles && threadIdx.x < 32) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_sg_tile(sidx ? stage0 : stage1, &full_bar[sidx ^ 1], s, maps, tn, tiles_per_img, PT);
        }
        mbar_wait(&full_bar[sidx], (it >> 1) & 1);
        const size_t b = t / tiles_per_img, pix0 = (size_t)(t % tiles_per_img) * PT;
        const bool active = pix0 + q < plane;
        const size_t pix = active ? pix0 + q : pix0;   // inactive lanes compute on pixel 0 of the tile, stores predicated
        // ---- this lane's 6 lobes (raw values from the stage) ----
        RenderSgArgs sv{};
        sv.axis = cur;
        sv.lamb = cur + 3 * K * PT;
        sv.weight = cur + 4 * K * PT;
        sv.K = K;
        sv.warp = s.warp;
        Lobes<float, kLPT> L;
        load_lobes(L, sv, 0, (size_t)q, (size_t)PT, h * kLPT);
        float W[kLPT][3];
#pragma unroll
        for (int k = 0; k < kLPT; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) W[k][c] = (h * kLPT + k < K) ? load_weight<float>(sv, 0, (size_t)q, (size_t)PT, h * kLPT + k, c) : 0.f;
        LobesPk<kLPT> Lp;
        pack_lobes<kLPT>(Lp, L);
        pk2 W2[3][NP2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < NP2; ++i) W2[c][i] = pk(W[2 * i][c], W[2 * i + 1][c]);

        if (MODE == 0 || MODE == 2 || MODE == 3) {
            if (MODE == 0 && active && (s.lamb_w || s.weight_w)) {
#pragma unroll
                for (int k = 0; k < kLPT; ++k) {
                    const int kk = h * kLPT + k;
                    if (kk < K) {
                        if (s.lamb_w) s.lamb_w[(b * K + kk) * plane + pix] = L.lam[k] * Consts<float>::LN2;
                        if (s.weight_w)
#pragma unroll
                            for (int c = 0; c < 3; ++c) s.weight_w[(b * (3 * K) + 3 * kk + c) * plane + pix] = W[k][c];
                    }
                }
            }
            const float sg = (MODE != 0 && active) ? s.seg[b * plane + pix] : 0.f;
            const float coef = (MODE == 3) ? reconst_coef(s.stats + b * 4) : 1.f;
            float acc0 = 0.f, acc1 = 0.f;   // MODE 2: sum E G, sum E^2;  MODE 3: sum dlog^2
            for (int d0 = 0; d0 < D; d0 += 4) {
                float E[3][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float wx = tab.lx[d0 + u], wy = tab.ly[d0 + u], wz = tab.lz[d0 + u];
                    pk2 arg[NP2], e[NP2];
                    lobe_exps_pk<kLPT>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg, e);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        pk2 acc = mul2(W2[c][0], e[0]);
#pragma unroll
                        for (int i = 1; i < NP2; ++i) acc = fma2(W2[c][i], e[i], acc);
                        E[c][u] = grp_allreduce<TPP>(hsum2(acc));
                    }
                }
                // the TPP lanes of a group share the 16-byte accesses: channel c of this 4-direction block belongs to
                // lane (c + d0/4) % TPP
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (active && ((c + (d0 >> 2)) % TPP) == h) {
                        if (MODE == 0) {
                            st4(s.env + ((b * 3 + c) * plane + pix) * D + d0, E[c][0], E[c][1], E[c][2], E[c][3]);
                        } else if (sg != 0.f) {
                            const F4 g = ld4(s.env_gt + ((b * 3 + c) * plane + pix) * D + d0);
                            const float G[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (MODE == 2) {
                                    acc0 += E[c][u] * G[u];
                                    acc1 += E[c][u] * E[c][u];
                                } else {
                                    const float dl = fast_log(coef * E[c][u] + s.offset) - fast_log(G[u] + s.offset);
                                    acc0 += dl * dl;
                                }
                            }
                        }
                    }
                }
            }
            if (MODE != 0) {
                // per-pixel weights: the regression sees (E s)(G s) and (E s)^2, the loss sees dlog^2 * s
                double* st = s.stats + b * 4;
                if (MODE == 2) {
                    sg_block_add(acc0 * sg * sg, st + 0);
                    sg_block_add(acc1 * sg * sg, st + 1);
                    sg_block_add((h == 0) ? sg : 0.f, st + 3);
                } else {
                    sg_block_add(acc0 * sg, st + 2);
                }
            }
        } else {
            pk2 gW2[3][NP2], S2[NP2], N2[NP2], mx2[NP2], my2[NP2], mz2[NP2];
#pragma unroll
            for (int i = 0; i < NP2; ++i) {
                S2[i] = N2[i] = mx2[i] = my2[i] = mz2[i] = pk(0.f, 0.f);
                gW2[0][i] = gW2[1][i] = gW2[2][i] = pk(0.f, 0.f);
            }
            const float* gsrc = (MODE >= 4) ? s.env_gt : s.g_env;
            const float* gp0 = gsrc + ((b * 3 + 0) * plane + pix) * D;
            const float* gp1 = gsrc + ((b * 3 + 1) * plane + pix) * D;
            const float* gp2 = gsrc + ((b * 3 + 2) * plane + pix) * D;
            const float sg4 = (MODE >= 4 && active) ? s.seg[b * plane + pix] : 0.f;
            const float coef4 = (MODE >= 4) ? reconst_coef(s.stats + b * 4) : 1.f;
            const float gfac = (MODE >= 4) ? 2.f * coef4 * sg4 * (MODE == 4 ? s.g_num[0] : 1.f) : 0.f;
            float lacc = 0.f;
            F4 n0 = ld4(gp0), n1 = ld4(gp1), n2 = ld4(gp2);
            for (int d0 = 0; d0 < D; d0 += 4) {
                const F4 g0 = n0, g1 = n1, g2 = n2;
                if (d0 + 4 < D) {   // next block of the stream, one block ahead of its use
                    n0 = ld4(gp0 + d0 + 4);
                    n1 = ld4(gp1 + d0 + 4);
                    n2 = ld4(gp2 + d0 + 4);
                }
                float G[3][4] = {{g0.x, g0.y, g0.z, g0.w}, {g1.x, g1.y, g1.z, g1.w}, {g2.x, g2.y, g2.z, g2.w}};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float wx = tab.lx[d0 + u], wy = tab.ly[d0 + u], wz = tab.lz[d0 + u];
                    pk2 arg[NP2], e[NP2];
                    lobe_exps_pk<kLPT>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg, e);
                    if (MODE >= 4) {   // dL/dE_cd = g * 2 c s (log(cE+off) - log(G+off)) / (cE+off)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            pk2 acc = mul2(W2[c][0], e[0]);
#pragma unroll
                            for (int i = 1; i < NP2; ++i) acc = fma2(W2[c][i], e[i], acc);
                            const float Ecd = grp_allreduce<TPP>(hsum2(acc));
                            const float ce = coef4 * Ecd + s.offset;
                            const float dl = fast_log(ce) - fast_log(G[c][u] + s.offset);
                            if (MODE == 5) lacc = fmaf(dl, dl, lacc);
                            G[c][u] = (sg4 != 0.f) ? gfac * dl * x_rcp(ce) : 0.f;   // masked-out pixel: its ground truth is never looked at
                        }
                    }
                    const pk2 G0 = pk(G[0][u], G[0][u]), G1 = pk(G[1][u], G[1][u]), G2 = pk(G[2][u], G[2][u]);
                    const pk2 wx2 = pk(wx, wx), wy2 = pk(wy, wy), wz2 = pk(wz, wz);
#pragma unroll
                    for (int i = 0; i < NP2; ++i) {
                        gW2[0][i] = fma2(G0, e[i], gW2[0][i]);
                        gW2[1][i] = fma2(G1, e[i], gW2[1][i]);
                        gW2[2][i] = fma2(G2, e[i], gW2[2][i]);
                        const pk2 tt = mul2(e[i], fma2(G0, W2[0][i], fma2(G1, W2[1][i], mul2(G2, W2[2][i]))));
                        S2[i] = fma2(tt, arg[i], S2[i]);
                        mx2[i] = fma2(tt, wx2, mx2[i]);
                        my2[i] = fma2(tt, wy2, my2[i]);
                        mz2[i] = fma2(tt, wz2, mz2[i]);
                    }
                }
            }
            // N_k = sum_d tt_kd = sum_c W_kc gW_kc (gW_kc = sum_d G_cd e_kd): formed once here instead of one packed add per
            // lobe pair and direction (1 of the loop's 15 FMA-pipe instructions)
#pragma unroll
            for (int i = 0; i < NP2; ++i) N2[i] = fma2(W2[0][i], gW2[0][i], fma2(W2[1][i], gW2[1][i], mul2(W2[2][i], gW2[2][i])));
            // every lane of a group holds the same sum; lanes past the end of the plane worked on unstaged data
            if (MODE == 5) sg_block_add((active && h == 0 && sg4 != 0.f) ? lacc * sg4 : 0.f, s.stats + b * 4 + 2);
            if (active) {
#pragma unroll
                for (int i = 0; i < NP2; ++i) {
                    float Sv[2], Nv[2], mxv[2], myv[2], mzv[2], gWv[3][2];
                    upk(S2[i], Sv[0], Sv[1]);
                    upk(N2[i], Nv[0], Nv[1]);
                    upk(mx2[i], mxv[0], mxv[1]);
                    upk(my2[i], myv[0], myv[1]);
                    upk(mz2[i], mzv[0], mzv[1]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) upk(gW2[c][i], gWv[c][0], gWv[c][1]);
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const int k = 2 * i + e2, kk = h * kLPT + k;
                        if (kk < K) {
                            const float lam = L.lam[k] * Consts<float>::LN2;
                            const size_t oa = ((b * K + kk) * 3) * plane + pix;
                            if (s.g_axis) {
                                s.g_axis[oa] = lam * mxv[e2];
                                s.g_axis[oa + plane] = lam * myv[e2];
                                s.g_axis[oa + 2 * plane] = lam * mzv[e2];
                            }
                            if (s.g_lamb) {
                                const size_t ia = ((size_t)kk * 3) * PT + q;
                                float g = lobe_lam_grad<float>(Sv[e2], L.lam[k], mxv[e2], myv[e2], mzv[e2], Nv[e2], sv.axis, ia, (size_t)PT);
                                if (s.warp) g *= sg_warp_grad<float>(lam);
                                s.g_lamb[(b * K + kk) * plane + pix] = g;
                            }
                            if (s.g_weight) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    float g = gWv[c][e2];
                                    if (s.warp) g *= sg_warp_grad<float>(W[k][c]);
                                    s.g_weight[(b * (3 * K) + 3 * kk + c) * plane + pix] = g;
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace irb
