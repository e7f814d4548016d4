// irb_render_sg_group.cuh -- LOBE-SPLIT persistent kernels of the fused SG render (device only).
//
// Why: with all K lobes of a pixel in one thread the backward needs ~250 registers, i.e. 2 warps per
// scheduler, and ncu shows the sweep then stalls ~45 % of its issue slots on fixed-latency dependencies
// (profiles/r01).  Here TPP = K/6 adjacent lanes share one pixel: each lane owns LPT = 6 lobes (72
// accumulator/parameter registers instead of 144) and every TPP-th direction's BRDF chain, so the
// register budget drops to <=128 (4 CTAs = 16 warps per SM) without duplicating any work:
//   * direction d = m*TPP + j: lane j of the group evaluates dir_fwd_local / dir_bwd_local for it and
//     broadcasts spec_d with one warp shuffle; every lane runs its 6 lobes for all directions;
//   * the per-direction  sum_k Q_k e_kd  is completed across the group with a shuffle reduce-scatter;
//   * per-pixel results (diffuse/spec, d/dnormal, d/drough, d/dalbedo) are butterfly-reduced once per pixel.
// Staging: persistent CTAs, 2-stage ring filled by tiled TMA (one cp.async.bulk.tensor per tensor) with full/empty
// mbarriers, tile = 128/TPP pixels.  The lobe-split fast path covers pixels in the regular regime (irb_math.cuh); the
// (rare) clamped / degenerate ones are shaded on the spot by the whole warp through the general world-space path of
// irb_driver.cuh (group_general_pixel below).
#pragma once

// A/B knobs (measured on B200, profiles/r02): rotating the backward loop so that the next pair's forward chain sits beside the
// current pair's backward chain, and pipelining the forward chain one pair ahead, were both 2-3 % SLOWER than ptxas' own
// schedule of the plain loops (bwd 1.926 vs 1.883 ms, fwd 0.854 vs 0.826 ms) -> off.
#ifndef IRB_BWD_ROTATE
#define IRB_BWD_ROTATE 0
#endif
#ifndef IRB_FWD_PIPE
#define IRB_FWD_PIPE 0
#endif

namespace irb {

// uv: the backward also stages the 2K planes of the forward's lobe sums (RenderSgArgs::lobe_uv_in) behind the other planes
__device__ __forceinline__ void issue_tile_n(float* stage, uint64_t* bar, const RenderSgArgs& a, const TileMaps& maps, size_t tile,
                                             size_t tiles_per_img, bool bwd, int PT, bool uv = false) {
    const int lane = threadIdx.x & 31;
    const size_t plane = (size_t)a.R * a.C;
    const size_t b = tile / tiles_per_img, pix0 = (tile % tiles_per_img) * (size_t)PT;
    if (maps.on) {
        // one tiled copy per tensor (7 forward / 9 backward) instead of one bulk copy per plane (7K+10 / 7K+16): the
        // per-plane form cost the producer warp ~75 cycles per copy, 11 % of its time, and the other warps waited for it
        if (lane == 0) {
            const int K = a.K;
            mbar_expect_tx(bar, (uint32_t)(stage_planes(K, bwd) + (uv ? 2 * K : 0)) * (uint32_t)PT * 4u);
            const int c0 = (int)pix0, cb = (int)b;
            tma_load_3d(stage, &maps.m[0], c0, 0, cb, bar);
            tma_load_3d(stage + 3 * PT, &maps.m[1], c0, 0, cb, bar);
            tma_load_3d(stage + 6 * PT, &maps.m[2], c0, 0, cb, bar);
            tma_load_3d(stage + 7 * PT, &maps.m[3], c0, 0, 0, bar);
            tma_load_3d(stage + 10 * PT, &maps.m[4], c0, 0, cb, bar);
            tma_load_3d(stage + (10 + 3 * K) * PT, &maps.m[5], c0, 0, cb, bar);
            tma_load_3d(stage + (10 + 4 * K) * PT, &maps.m[6], c0, 0, cb, bar);
            if (bwd) {
                tma_load_3d(stage + (10 + 7 * K) * PT, &maps.m[7], c0, 0, cb, bar);
                tma_load_3d(stage + (13 + 7 * K) * PT, &maps.m[8], c0, 0, cb, bar);
                if (uv) tma_load_3d(stage + (16 + 7 * K) * PT, &maps.m[9], c0, 0, cb, bar);
            }
        }
        return;
    }
    const uint32_t valid = (uint32_t)((plane - pix0) < (size_t)PT ? (plane - pix0) : (size_t)PT);
    const uint32_t bytes = valid * 4u;
    const int K = a.K;
    const int np = stage_planes(K, bwd) + (uv ? 2 * K : 0);
    if (lane == 0) mbar_expect_tx(bar, bytes * (uint32_t)np);
    __syncwarp();
    for (int p = lane; p < np; p += 32) {
        const float* src;
        if (p < 3) src = a.albedo + (b * 3 + p) * plane;
        else if (p < 6) src = a.normal + (b * 3 + (p - 3)) * plane;
        else if (p < 7) src = a.rough + b * plane;
        else if (p < 10) src = a.view + (size_t)(p - 7) * plane;
        else if (p < 10 + 3 * K) src = a.axis + (b * 3 * K + (p - 10)) * plane;
        else if (p < 10 + 4 * K) src = a.lamb + (b * K + (p - 10 - 3 * K)) * plane;
        else if (p < 10 + 7 * K) src = a.weight + (b * 3 * K + (p - 10 - 4 * K)) * plane;
        else if (p < 13 + 7 * K) src = a.g_diff + (b * 3 + (p - 10 - 7 * K)) * plane;
        else if (p < 16 + 7 * K) src = a.g_spec + (b * 3 + (p - 13 - 7 * K)) * plane;
        else src = a.lobe_uv_in + (b * 2 * K + (p - 16 - 7 * K)) * plane;
        bulk_g2s(stage + (size_t)p * PT, src + pix0, bytes, bar);
    }
}

__device__ __forceinline__ RenderSgArgs stage_view_n(const RenderSgArgs& a, float* stage, int PT) {
    RenderSgArgs s = a;
    const int K = a.K;
    s.albedo = stage;
    s.normal = stage + 3 * PT;
    s.rough = stage + 6 * PT;
    s.view = stage + 7 * PT;
    s.axis = stage + 10 * PT;
    s.lamb = stage + (10 + 3 * K) * PT;
    s.weight = stage + (10 + 4 * K) * PT;
    s.g_diff = stage + (10 + 7 * K) * PT;
    s.g_spec = stage + (13 + 7 * K) * PT;
    s.lobe_uv_in = stage + (16 + 7 * K) * PT;
    return s;
}

// hemisphere table in shared memory (the BRDF direction index differs between the lanes of a group)
struct TabS {
    float lx[IRB_MAX_DIRS], ly[IRB_MAX_DIRS], lz[IRB_MAX_DIRS], w[IRB_MAX_DIRS];
};


// ---- one tile, forward ----------------------------------------------------------------------
template <int TPP>
__device__ __forceinline__ bool group_tile_fwd(const RenderSgArgs& a, const RenderSgArgs& sv, const EnvTab& tab, const TabS& ts,
                                               const PixView& vi, const PixView& vo, bool in_range, int h) {
    const int lane = threadIdx.x & 31, lane_base = lane - h;
    Pixel<float> px;
    V3<float> nt;
    float A[3];
    load_pixel(px, nt, A, sv, vi.b, vi.pix, vi.plane);
    const bool active = in_range && px.regular;
    PixelLocal<float> pl;
    pixel_local_setup(pl, px);
    Lobes<float, kLPT> L;
    load_lobes(L, sv, vi.b, vi.pix, vi.plane, h * kLPT);
    float U[kLPT], V[kLPT];
    LobesPk<kLPT> Lp;
    pack_lobes<kLPT>(Lp, L);
    pk2 U2[kLPT / 2], V2[kLPT / 2];
#pragma unroll
    for (int i = 0; i < kLPT / 2; ++i) U2[i] = V2[i] = pk(0.f, 0.f);
    // 2*TPP directions per iteration: this lane evaluates the BRDF chains of directions m+h and m+TPP+h packed in
    // the two halves of f32x2 registers
#if IRB_FWD_PIPE
    // software pipeline: the BRDF chain of the NEXT direction pair is issued in the same basic block as the lobe work of
    // the current one (the chain is ~25 dependent steps deep; alone it leaves the FMA / XU pipes idle)
    DirL2 dr;
    dir_fwd_local2(dr, px, pl, pk(ts.lx[h], ts.lx[TPP + h]), pk(ts.ly[h], ts.ly[TPP + h]), pk(ts.lz[h], ts.lz[TPP + h]));
    for (int m = 0; m < tab.D; m += 2 * TPP) {
        float sp_lo, sp_hi;
        upk(dr.spec, sp_lo, sp_hi);
        {
            const int mn = (m + 2 * TPP < tab.D) ? m + 2 * TPP : m;   // clamped on the last trip (result unused)
            const int d0 = mn + h, d1 = mn + TPP + h;
            dir_fwd_local2(dr, px, pl, pk(ts.lx[d0], ts.lx[d1]), pk(ts.ly[d0], ts.ly[d1]), pk(ts.lz[d0], ts.lz[d1]));
        }
#else
    for (int m = 0; m < tab.D; m += 2 * TPP) {
        const int d0 = m + h, d1 = m + TPP + h;
        DirL2 dr;
        dir_fwd_local2(dr, px, pl, pk(ts.lx[d0], ts.lx[d1]), pk(ts.ly[d0], ts.ly[d1]), pk(ts.lz[d0], ts.lz[d1]));
        float sp_lo, sp_hi;
        upk(dr.spec, sp_lo, sp_hi);
#endif
        pk2 e[2 * TPP][kLPT / 2];
        pk2 w2[2 * TPP], ws2[2 * TPP];
#pragma unroll
        for (int j = 0; j < 2 * TPP; ++j) {
            const int d = m + j;
            const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
            const float spec = __shfl_sync(0xffffffffu, j < TPP ? sp_lo : sp_hi, lane_base + (j % TPP));
            const float w = tab.w[d] * wz;
            const float ws = w * spec;
            pk2 arg[kLPT / 2];
            lobe_exps_pk<kLPT>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg, e[j]);
            w2[j] = pk(w, w);
            ws2[j] = pk(ws, ws);
        }
#pragma unroll
        for (int j = 0; j < 2 * TPP; ++j) {
#pragma unroll
            for (int i = 0; i < kLPT / 2; ++i) {
                U2[i] = fma2(w2[j], e[j][i], U2[i]);
                V2[i] = fma2(ws2[j], e[j][i], V2[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kLPT / 2; ++i) {
        upk(U2[i], U[2 * i], U[2 * i + 1]);
        upk(V2[i], V[2 * i], V[2 * i + 1]);
    }
    if (a.lobe_uv_out && active) {   // the backward's U_k, V_k (regular pixels; the others recompute them in the general path)
#pragma unroll
        for (int k = 0; k < kLPT; ++k) {
            const int kk = h * kLPT + k;
            if (kk < sv.K) {
                float* o = a.lobe_uv_out + (vo.b * (2 * sv.K) + 2 * kk) * vo.plane + vo.pix;
                o[0] = U[k];
                o[vo.plane] = V[k];
            }
        }
    }
    float Dc[3] = {0.f, 0.f, 0.f}, Sc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kLPT; ++k) {
        if (h * kLPT + k < sv.K) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float w = load_weight<float>(sv, vi.b, vi.pix, vi.plane, h * kLPT + k, c);
                Dc[c] += w * U[k];
                Sc[c] += w * V[k];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Dc[c] = grp_allreduce<TPP>(Dc[c]);
        Sc[c] = grp_allreduce<TPP>(Sc[c]);
    }
    if (active && h == 0) {
        const size_t o3 = (vo.b * 3) * vo.plane + vo.pix;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a.out_diff[o3 + c * vo.plane] = A[c] * Consts<float>::INV_PI * Dc[c];
            a.out_spec[o3 + c * vo.plane] = Sc[c];
        }
    }
    return in_range && !px.regular;
}

// ---- one tile, backward ---------------------------------------------------------------------
template <int TPP, bool PIX, bool LOBE, bool SAVED>
__device__ __forceinline__ bool group_tile_bwd(const RenderSgArgs& a, const RenderSgArgs& sv, const EnvTab& tab, const TabS& ts,
                                               const PixView& vi, const PixView& vo, bool in_range, int h) {
    const int lane = threadIdx.x & 31, lane_base = lane - h;
    const size_t i3 = (vi.b * 3) * vi.plane + vi.pix;
    float gD[3], gS[3], gDa[3];
    bool active, irregular;
    PixelGradL<float> gl;
    pixel_gradl_zero(gl);
    LobeGrads<float, kLPT> lg;
    Lobes<float, kLPT> L;
    float P[kLPT], Q[kLPT];
    {
        Pixel<float> px;
        V3<float> nt;
        float A[3];
        load_pixel(px, nt, A, sv, vi.b, vi.pix, vi.plane);
        active = in_range && px.regular;
        irregular = in_range && !px.regular;
        PixelLocal<float> pl;
        pixel_local_setup(pl, px);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gD[c] = sv.g_diff[i3 + c * vi.plane];
            gS[c] = sv.g_spec[i3 + c * vi.plane];
            gDa[c] = gD[c] * A[c] * Consts<float>::INV_PI;
        }
        load_lobes(L, sv, vi.b, vi.pix, vi.plane, h * kLPT);
        // P, Q from the warped intensities; the warped values overwrite the raw ones in this thread's own slots of the
        // stage so that the epilogue (chain rule, d/dalbedo) does not evaluate the tan warp a second time
        {
            float* wst = const_cast<float*>(sv.weight);
#pragma unroll
            for (int k = 0; k < kLPT; ++k) {
                P[k] = Q[k] = 0.f;
                const int kk = h * kLPT + k;
                if (kk < sv.K) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float w = load_weight<float>(sv, vi.b, vi.pix, vi.plane, kk, c);
                        wst[(vi.b * (3 * sv.K) + 3 * kk + c) * vi.plane + vi.pix] = w;
                        P[k] += gDa[c] * w;
                        Q[k] += gS[c] * w;
                    }
                }
            }
        }
        LobesPk<kLPT> Lp;
        pack_lobes<kLPT>(Lp, L);
        pk2 P2[kLPT / 2], Q2[kLPT / 2];
        pack_pq<kLPT>(P2, Q2, P, Q);
        LobeGradsPk<kLPT> g2;
        gradspk_zero<kLPT>(g2);
        PixelGradL2 gl2;
        pixel_gradl2_zero(gl2);
#if IRB_BWD_ROTATE
        // Rotated loop: the BRDF forward chain of the NEXT direction pair sits next to the BRDF backward chain of the
        // current one (two independent, long, serially dependent chains in one basic block), so that the scheduler can
        // interleave them -- with 2 warps per scheduler (255 registers) a lone chain leaves the FMA pipe idle for most of
        // its ~25-deep dependency path (ncu r01: stall_wait 32 %).  The lobe temporaries are dead in that phase, so the
        // second DirL2 does not raise the register peak.  Unrolled by two with ping-pong state (no register copies).
        struct DirState {
            DirL2 dr;
            pk2 wx, wy, wz;
        };
        auto load_dirs = [&](DirState& st, int m) {
            st.wx = pk(ts.lx[m + h], ts.lx[m + TPP + h]);
            st.wy = pk(ts.ly[m + h], ts.ly[m + TPP + h]);
            st.wz = pk(ts.lz[m + h], ts.lz[m + TPP + h]);
            dir_fwd_local2(st.dr, px, pl, st.wx, st.wy, st.wz);
        };
        auto body = [&](int m, DirState& cur, DirState& nxt) {
            float sp_lo, sp_hi;
            upk(cur.dr.spec, sp_lo, sp_hi);
            float part[2 * TPP];
#pragma unroll
            for (int half = 0; half < 2; ++half) {   // the two direction groups in turn: keeps only TPP*3 (arg, e) pairs live
                pk2 arg[TPP][kLPT / 2], e[TPP][kLPT / 2];
                float spec[TPP];
#pragma unroll
                for (int j = 0; j < TPP; ++j) {   // phase 1: arguments + exponentials
                    const int d = m + half * TPP + j;
                    const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                    spec[j] = __shfl_sync(0xffffffffu, half ? sp_hi : sp_lo, lane_base + j);
                    lobe_exps_pk<kLPT>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg[j], e[j]);
                }
#pragma unroll
                for (int j = 0; j < TPP; ++j) {   // phase 2: accumulations
                    const int d = m + half * TPP + j;
                    const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                    part[half * TPP + j] =
                        lobes_bwd_acc_pk<kLPT, PIX, LOBE, SAVED>(g2, P2, Q2, arg[j], e[j], wx, wy, wz, tab.w[d] * wz, spec[j]);
                }
            }
            // next pair (clamped on the last trip: recomputes the last pair, result unused)
            load_dirs(nxt, (m + 2 * TPP < tab.D) ? m + 2 * TPP : m);
            if (PIX) {
                float pa[TPP], pb[TPP];
#pragma unroll
                for (int j = 0; j < TPP; ++j) {
                    pa[j] = part[j];
                    pb[j] = part[TPP + j];
                }
                const float s0 = grp_reduce_scatter<TPP>(pa, h), s1 = grp_reduce_scatter<TPP>(pb, h);
                float z0, z1;
                upk(cur.wz, z0, z1);
                dir_bwd_local2(gl2, px, pl, cur.dr, cur.wx, cur.wy, cur.wz,
                               pk(ts.w[m + h] * z0 * s0, ts.w[m + TPP + h] * z1 * s1));
            }
        };
        DirState sa, sb;
        load_dirs(sa, 0);
        int m = 0;
        for (; m + 4 * TPP <= tab.D; m += 4 * TPP) {
            body(m, sa, sb);
            body(m + 2 * TPP, sb, sa);
        }
        if (m < tab.D) body(m, sa, sb);
#else
        for (int m = 0; m < tab.D; m += 2 * TPP) {
            const int d0 = m + h, d1 = m + TPP + h;
            const pk2 wxm = pk(ts.lx[d0], ts.lx[d1]), wym = pk(ts.ly[d0], ts.ly[d1]), wzm = pk(ts.lz[d0], ts.lz[d1]);
            DirL2 dr;
            dir_fwd_local2(dr, px, pl, wxm, wym, wzm);
            float sp_lo, sp_hi;
            upk(dr.spec, sp_lo, sp_hi);
            float part[2 * TPP];
#pragma unroll
            for (int half = 0; half < 2; ++half) {   // the two direction groups in turn: keeps only TPP*3 (arg, e) pairs live
                pk2 arg[TPP][kLPT / 2], e[TPP][kLPT / 2];
                float spec[TPP];
#pragma unroll
                for (int j = 0; j < TPP; ++j) {   // phase 1: arguments + exponentials
                    const int d = m + half * TPP + j;
                    const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                    spec[j] = __shfl_sync(0xffffffffu, half ? sp_hi : sp_lo, lane_base + j);
                    lobe_exps_pk<kLPT>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg[j], e[j]);
                }
#pragma unroll
                for (int j = 0; j < TPP; ++j) {   // phase 2: accumulations
                    const int d = m + half * TPP + j;
                    const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                    part[half * TPP + j] =
                        lobes_bwd_acc_pk<kLPT, PIX, LOBE, SAVED>(g2, P2, Q2, arg[j], e[j], wx, wy, wz, tab.w[d] * wz, spec[j]);
                }
            }
            if (PIX) {
                float pa[TPP], pb[TPP];
#pragma unroll
                for (int j = 0; j < TPP; ++j) {
                    pa[j] = part[j];
                    pb[j] = part[TPP + j];
                }
                const float s0 = grp_reduce_scatter<TPP>(pa, h), s1 = grp_reduce_scatter<TPP>(pb, h);
                float z0, z1;
                upk(wzm, z0, z1);
                dir_bwd_local2(gl2, px, pl, dr, wxm, wym, wzm, pk(ts.w[d0] * z0 * s0, ts.w[d1] * z1 * s1));
            }
        }
#endif
        pixel_gradl2_fold(gl, gl2);
        gradspk_unpack<kLPT>(lg, g2);
        if (SAVED) {   // U_k, V_k of the forward pass, staged with the tile
#pragma unroll
            for (int k = 0; k < kLPT; ++k) {
                const int kk = h * kLPT + k;
                const bool okk = kk < sv.K;
                lg.U[k] = okk ? sv.lobe_uv_in[(size_t)(2 * kk) * vi.plane + vi.pix] : 0.f;
                lg.V[k] = okk ? sv.lobe_uv_in[(size_t)(2 * kk + 1) * vi.plane + vi.pix] : 0.f;
            }
        }
    }
    // epilogue: outputs are predicated per thread by nulling the pointers of inactive threads
    RenderSgArgs ao = a;
    if (!active) ao.g_axis = ao.g_lamb = ao.g_weight = nullptr;
    float gAlb[3] = {0.f, 0.f, 0.f};
    RenderSgArgs svw = sv;   // stage view whose weight planes now hold WARPED values
    svw.warp = 0;
    if (LOBE) store_lobe_grads<float, kLPT>(lg, L, P, Q, svw, vi, ao, vo, h * kLPT, gDa, gS, gAlb);
    const size_t o3 = (vo.b * 3) * vo.plane + vo.pix;
    if (LOBE && a.g_albedo) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gAlb[c] = grp_allreduce<TPP>(gAlb[c]);
        if (active && h == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) a.g_albedo[o3 + c * vo.plane] = gD[c] * Consts<float>::INV_PI * gAlb[c];
        }
    }
    if (PIX) {
        gl.Gxx = grp_allreduce<TPP>(gl.Gxx); gl.Gxy = grp_allreduce<TPP>(gl.Gxy); gl.Gxz = grp_allreduce<TPP>(gl.Gxz);
        gl.Gyx = grp_allreduce<TPP>(gl.Gyx); gl.Gyy = grp_allreduce<TPP>(gl.Gyy); gl.Gyz = grp_allreduce<TPP>(gl.Gyz);
        gl.Gzx = grp_allreduce<TPP>(gl.Gzx); gl.Gzy = grp_allreduce<TPP>(gl.Gzy); gl.Gzz = grp_allreduce<TPP>(gl.Gzz);
        gl.gnx = grp_allreduce<TPP>(gl.gnx); gl.gny = grp_allreduce<TPP>(gl.gny); gl.gnz = grp_allreduce<TPP>(gl.gnz);
        gl.G_n1 = grp_allreduce<TPP>(gl.G_n1);
        gl.g_alpha2 = grp_allreduce<TPP>(gl.g_alpha2);
        gl.g_k = grp_allreduce<TPP>(gl.g_k);
        if (active && h == 0) {
            // the world-space pixel state is recomputed here instead of being kept alive through the sweep
            Pixel<float> px;
            V3<float> nt;
            float A[3];
            load_pixel(px, nt, A, sv, vi.b, vi.pix, vi.plane);
            PixelGrad<float> pg;
            gradl_to_world(pg, gl, px);
            V3<float> g_nt;
            float g_rough;
            pixel_bwd_finish(pg, px, nt, g_nt, g_rough);
            if (a.g_normal) {
                a.g_normal[o3] = g_nt.x;
                a.g_normal[o3 + vo.plane] = g_nt.y;
                a.g_normal[o3 + 2 * vo.plane] = g_nt.z;
            }
            if (a.g_rough) a.g_rough[vo.b * vo.plane + vo.pix] = g_rough;
        }
    }
    return irregular;
}

// ---- pixels outside the regular regime (clamped / degenerate normals) ------------------------------------------------
// They take the general world-space path of irb_driver.cuh -- the code of the plain kernels.  Real inputs have none or a
// handful (pooled normals that happen to be parallel to the up vector), but one thread walking all 128 directions x K
// lobes alone takes ~60-100 us, which a follow-up launch (or a queue drained at the end of the kernel) exposes in full.
// So the warp that meets such a pixel shades it at once, cooperatively: every lane takes 4 of the 128 directions, the
// partial sums are added across the warp, lane 0 stores (a few microseconds, inside a kernel that runs for a millisecond).
// __noinline__: the general path gets its own register allocation instead of inflating the fast path's.
struct WarpCoop {
    __device__ static DirSlice slice() { return DirSlice{(int)(threadIdx.x & 31), 32}; }
    __device__ static void reduce(float& v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    __device__ static bool leader() { return (threadIdx.x & 31) == 0; }
};

template <int MODE, bool PIX, bool LOBE>
__device__ __noinline__ void group_general_pixel(const RenderSgArgs& a, const EnvTab& tab, unsigned gidx) {
    const size_t plane = (size_t)a.R * a.C;
    const PixView v{(size_t)gidx / plane, (size_t)gidx % plane, plane};
    if (MODE == 0) {
        if (a.K <= 12) drv_render_sg_fwd<float, 12, WarpCoop>(a, v, a, v, tab);
        else drv_render_sg_fwd<float, 24, WarpCoop>(a, v, a, v, tab);
    } else {
        if (a.K <= 12) drv_render_sg_bwd<float, 12, 12, PIX, LOBE, WarpCoop>(a, v, a, v, tab);
        else drv_render_sg_bwd<float, 24, 12, PIX, LOBE, WarpCoop>(a, v, a, v, tab);
    }
}

// MODE 0 forward, 1 backward
// FULLK: K == TPP * kLPT exactly (6 / 12 / 24 lobes): the per-lobe `kk < K` guards, the padding lobes and the runtime-K address
// arithmetic of the once-per-pixel prologue / epilogue fold away (that code is ~20 % of the instructions a tile executes)
template <int MODE, int TPP, bool PIX, bool LOBE, int MINB, bool SAVED = false, bool FULLK = false>
__global__ void __launch_bounds__(kThreads, MINB) render_sg_group_kernel(const __grid_constant__ RenderSgArgs a, const __grid_constant__ EnvTab tab,
                                                                          const __grid_constant__ TileMaps maps,
                                                                          const unsigned total_tiles, const unsigned tiles_per_img) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
    __shared__ TabS ts;
    constexpr bool BWD = (MODE == 1);
    constexpr int PT = kThreads / TPP;
    const int np = stage_planes(a.K, BWD) + (SAVED ? 2 * a.K : 0);
    float* const stage0 = reinterpret_cast<float*>(smem_raw);
    float* const stage1 = stage0 + (size_t)np * PT;
    for (int i = threadIdx.x; i < IRB_MAX_DIRS; i += kThreads) {
        ts.lx[i] = tab.lx[i];
        ts.ly[i] = tab.ly[i];
        ts.lz[i] = tab.lz[i];
        ts.w[i] = tab.w[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_init(&empty_bar[0], kThreads);
        mbar_init(&empty_bar[1], kThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Producer / consumer ring without block-wide barriers: warp 0 refills a stage as soon as all 128 threads have
    // released it (empty barrier), the other warps never wait for each other, so the low-IPC per-tile prologue and
    // epilogue of one warp overlap the direction sweeps of the others (ncu: 11 % barrier stalls with __syncthreads).
    const size_t plane = (size_t)a.R * a.C;
    const int q = threadIdx.x / TPP, h = threadIdx.x % TPP;
    unsigned t = blockIdx.x;
    if (t < total_tiles && threadIdx.x < 32) issue_tile_n(stage0, &full_bar[0], a, maps, t, tiles_per_img, BWD, PT, SAVED);
    for (unsigned it = 0; t < total_tiles; ++it, t += gridDim.x) {
        const int s = it & 1;
        float* const cur = s ? stage1 : stage0;
        const unsigned tn = t + gridDim.x;
        if (tn < total_tiles && threadIdx.x < 32) {
            const unsigned kfill = (it + 1) >> 1;            // fill index of stage s^1
            if (kfill >= 1) mbar_wait(&empty_bar[s ^ 1], (kfill - 1) & 1);   // everyone finished the previous tenant
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile_n(s ? stage0 : stage1, &full_bar[s ^ 1], a, maps, tn, tiles_per_img, BWD, PT, SAVED);
        }
        mbar_wait(&full_bar[s], (it >> 1) & 1);
        const size_t b = t / tiles_per_img, pix0 = (size_t)(t % tiles_per_img) * PT;
        const bool in_range = pix0 + q < plane;
        RenderSgArgs sv = stage_view_n(a, cur, PT);
        RenderSgArgs ak = a;
        if (FULLK) ak.K = sv.K = TPP * kLPT;   // compile-time lobe count
        const PixView vi{0, (size_t)q, (size_t)PT};
        const PixView vo{b, pix0 + q, plane};
        bool irr;
        if (MODE == 0) irr = group_tile_fwd<TPP>(ak, sv, tab, ts, vi, vo, in_range, h);
        else irr = group_tile_bwd<TPP, PIX, LOBE, SAVED>(ak, sv, tab, ts, vi, vo, in_range, h);
        mbar_arrive(&empty_bar[s]);   // this thread is done with stage s
        unsigned m = __ballot_sync(0xffffffffu, irr && h == 0);
        while (m) {   // rare: the whole warp shades each such pixel through the general path
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const unsigned g = __shfl_sync(0xffffffffu, (unsigned)(b * plane + pix0 + q), src);
            group_general_pixel<MODE, PIX, LOBE>(a, tab, g);
        }
    }
}

}  // namespace irb
