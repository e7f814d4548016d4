// irb_driver.cuh -- per-pixel drivers of the fused SG render: load one pixel's planes, run the
// register-resident sweeps of irb_pixel.cuh, store results.  One CUDA thread runs one driver
// call (thread-per-pixel: the lobe parameters, their gradient accumulators and the BRDF state
// all live in registers; channel-planar tensors make every load/store coalesced across the
// warp).  The same code is instantiated on the host by tests/hostcheck (T = float / double).
#pragma once
#include "irb_pixel.cuh"

namespace irb {

struct RenderSgArgs {
    // inputs (models.py:461 forwardEnv + models.py:391 output2env)
    const float* albedo;  // [B,3,R,C] pooled
    const float* normal;  // [B,3,R,C] pooled, un-normalised
    const float* rough;   // [B,1,R,C] pooled
    const float* axis;    // [B,K,3,R,C]
    const float* lamb;    // [B,K,R,C]
    const float* weight;  // [B,3K,R,C]  channel = 3k+c
    const float* view;    // [3,R,C]
    // forward outputs
    float* out_diff;  // [B,3,R,C]
    float* out_spec;
    // backward
    const float* g_diff;
    const float* g_spec;
    float* g_albedo;
    float* g_normal;
    float* g_rough;
    float* g_axis;
    float* g_lamb;
    float* g_weight;
    // optional [B,2K,R,C] side tensor: plane 2k = U_k = sum_d w_d e_kd, plane 2k+1 = V_k = sum_d w_d spec_d e_kd.
    // The forward writes it (regular pixels only), the backward then skips the two accumulations per lobe and direction that
    // would only rebuild these sums (10 % of its FMA-pipe work; HBM has 90 % slack on this path).  NULL = recompute.
    float* lobe_uv_out;
    const float* lobe_uv_in;
    int B, K, R, C;
    float F0;
    int warp;
};

template <typename T, int KC>
IRB_HD void load_lobes(Lobes<T, KC>& L, const RenderSgArgs& a, size_t b, size_t pix, size_t plane, int k0) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int kk = k0 + k;
        if (kk < a.K) {
            const size_t oa = ((b * a.K + kk) * 3) * plane + pix;
            T lam = (T)a.lamb[(b * a.K + kk) * plane + pix];
            if (a.warp) lam = sg_warp<T>(lam);
            lam *= Consts<T>::LOG2E;
            L.lam[k] = lam;
            L.bx[k] = lam * (T)a.axis[oa];
            L.by[k] = lam * (T)a.axis[oa + plane];
            L.bz[k] = lam * (T)a.axis[oa + 2 * plane];
        } else {  // padding lobe: contributes exactly zero (its weight is taken as 0)
            L.bx[k] = L.by[k] = L.bz[k] = T(0);
            L.lam[k] = T(0);
        }
    }
}

// dL/dlam_k = sum_d T (a.w_d - 1) = S / lam'   (S = sum_d T arg, arg = lam' (a.w_d - 1)).
// lam' == 0 (lambOrig == 0 exactly, a value the LightNet head can emit): e == 1 and
// sum_d T (a.w_d - 1) = a.M - N with N = sum_d T = P U + Q V.
template <typename T>
IRB_HD T lobe_lam_grad(T S, T lam, T mx, T my, T mz, T N, const float* axis, size_t oa, size_t plane) {
    if (lam != T(0)) return x_div(S, lam);
    return (T)axis[oa] * mx + (T)axis[oa + plane] * my + (T)axis[oa + 2 * plane] * mz - N;
}

template <typename T> IRB_HD T load_weight(const RenderSgArgs& a, size_t b, size_t pix, size_t plane, int kk, int c) {
    T w = (T)a.weight[(b * (3 * a.K) + 3 * kk + c) * plane + pix];
    if (a.warp) w = sg_warp<T>(w);
    return w;
}

template <typename T>
IRB_HD void load_pixel(Pixel<T>& px, V3<T>& nt, T (&A)[3], const RenderSgArgs& a, size_t b, size_t pix, size_t plane) {
    const size_t o3 = (b * 3) * plane + pix;
    nt = V3<T>{(T)a.normal[o3], (T)a.normal[o3 + plane], (T)a.normal[o3 + 2 * plane]};
    A[0] = (T)a.albedo[o3];
    A[1] = (T)a.albedo[o3 + plane];
    A[2] = (T)a.albedo[o3 + 2 * plane];
    const V3<T> v{(T)a.view[pix], (T)a.view[plane + pix], (T)a.view[2 * plane + pix]};
    pixel_setup(px, nt, (T)a.rough[b * plane + pix], v, (T)a.F0);
}

// A driver call reads its inputs through an INPUT VIEW (ai, bi, pi, pli) and writes through an OUTPUT VIEW
// (a, b, pix, plane).  The plain kernels and tests/hostcheck pass the same view twice (global tensors);
// the TMA-staged persistent kernels pass a view of the shared-memory stage as the input view.
struct PixView {
    size_t b, pix, plane;
};

// Who works on a pixel.  Solo: one thread owns it (plain kernels, host check).  A warp-cooperative policy
// (irb_render_sg_group.cuh) gives every lane a slice of the directions, sums the partial results across the warp and
// lets one lane store.
struct Solo {
    IRB_HD static DirSlice slice() { return DirSlice{0, 1}; }
    template <typename T> IRB_HD static void reduce(T&) {}
    IRB_HD static bool leader() { return true; }
};
template <typename C, typename T> IRB_HD void coop_reduce(PixelGrad<T>& g) {
    C::reduce(g.g_cx.x); C::reduce(g.g_cx.y); C::reduce(g.g_cx.z);
    C::reduce(g.g_cy.x); C::reduce(g.g_cy.y); C::reduce(g.g_cy.z);
    C::reduce(g.g_n.x); C::reduce(g.g_n.y); C::reduce(g.g_n.z);
    C::reduce(g.G_n1); C::reduce(g.g_alpha2); C::reduce(g.g_k);
}
template <typename C, typename T, int KC> IRB_HD void coop_reduce(LobeGrads<T, KC>& g) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        C::reduce(g.U[k]); C::reduce(g.V[k]); C::reduce(g.S[k]);
        C::reduce(g.mx[k]); C::reduce(g.my[k]); C::reduce(g.mz[k]);
    }
}

// ---- forward: out_diff, out_spec --------------------------------------------------------
template <typename T, int KC, typename Coop = Solo>
IRB_HD void drv_render_sg_fwd(const RenderSgArgs& ai, const PixView& vi, const RenderSgArgs& a, const PixView& vo, const EnvTab& tab) {
    Pixel<T> px;
    V3<T> nt;
    T A[3];
    load_pixel(px, nt, A, ai, vi.b, vi.pix, vi.plane);
    T Dc[3] = {0, 0, 0}, Sc[3] = {0, 0, 0};
    for (int k0 = 0; k0 < ai.K; k0 += KC) {
        Lobes<T, KC> L;
        load_lobes(L, ai, vi.b, vi.pix, vi.plane, k0);
        T U[KC], V[KC];
        sweep_fwd<T, KC>(px, L, tab, U, V, Coop::slice());
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            Coop::reduce(U[k]);
            Coop::reduce(V[k]);
            if (k0 + k < ai.K) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const T w = load_weight<T>(ai, vi.b, vi.pix, vi.plane, k0 + k, c);
                    Dc[c] += w * U[k];
                    Sc[c] += w * V[k];
                }
            }
        }
    }
    if (!Coop::leader()) return;
    const size_t o3 = (vo.b * 3) * vo.plane + vo.pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a.out_diff[o3 + c * vo.plane] = (float)(A[c] * Consts<T>::INV_PI * Dc[c]);
        a.out_spec[o3 + c * vo.plane] = (float)Sc[c];
    }
}
template <typename T, int KC>
IRB_HD void drv_render_sg_fwd(const RenderSgArgs& a, const EnvTab& tab, size_t b, size_t pix) {
    const PixView v{b, pix, (size_t)a.R * a.C};
    drv_render_sg_fwd<T, KC>(a, v, a, v, tab);
}

// ---- backward -----------------------------------------------------------------------------
// KP: chunk width of the BRDF-side sweep (must cover all K lobes); KL: chunk width of the
// SG-side sweeps.  When KP == KL and K <= KP a single sweep does both.
template <typename T, int KC>
IRB_HD void load_PQ(T (&P)[KC], T (&Q)[KC], const RenderSgArgs& a, size_t b, size_t pix, size_t plane, int k0,
                    const T (&gDa)[3], const T (&gS)[3]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        P[k] = Q[k] = T(0);
        if (k0 + k < a.K) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const T w = load_weight<T>(a, b, pix, plane, k0 + k, c);
                P[k] += gDa[c] * w;
                Q[k] += gS[c] * w;
            }
        }
    }
}

template <typename T, int KC>
IRB_HD void store_lobe_grads(const LobeGrads<T, KC>& lg, const Lobes<T, KC>& L, const T (&P)[KC], const T (&Q)[KC],
                             const RenderSgArgs& ai, const PixView& vi, const RenderSgArgs& a, const PixView& vo, int k0,
                             const T (&gDa)[3], const T (&gS)[3], T (&gAlb)[3]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int kk = k0 + k;
        if (kk < a.K) {
            const T lam = L.lam[k] * Consts<T>::LN2;  // undo the log2e fold
            const size_t oa = ((vo.b * a.K + kk) * 3) * vo.plane + vo.pix;
            if (a.g_axis) {
                a.g_axis[oa] = (float)(lam * lg.mx[k]);
                a.g_axis[oa + vo.plane] = (float)(lam * lg.my[k]);
                a.g_axis[oa + 2 * vo.plane] = (float)(lam * lg.mz[k]);
            }
            if (a.g_lamb) {
                const size_t ia = ((vi.b * ai.K + kk) * 3) * vi.plane + vi.pix;
                T g = lobe_lam_grad<T>(lg.S[k], L.lam[k], lg.mx[k], lg.my[k], lg.mz[k], P[k] * lg.U[k] + Q[k] * lg.V[k],
                                       ai.axis, ia, vi.plane);
                if (a.warp) g *= sg_warp_grad<T>(lam);
                a.g_lamb[(vo.b * a.K + kk) * vo.plane + vo.pix] = (float)g;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const T w = load_weight<T>(ai, vi.b, vi.pix, vi.plane, kk, c);
                gAlb[c] += w * lg.U[k];
                if (a.g_weight) {
                    T g = gDa[c] * lg.U[k] + gS[c] * lg.V[k];
                    if (a.warp) g *= sg_warp_grad<T>(w);
                    a.g_weight[(vo.b * (3 * a.K) + 3 * kk + c) * vo.plane + vo.pix] = (float)g;
                }
            }
        }
    }
}

template <typename T, int KP, int KL, bool PIX, bool LOBE, typename Coop = Solo>
IRB_HD void drv_render_sg_bwd(const RenderSgArgs& ai, const PixView& vi, const RenderSgArgs& a, const PixView& vo, const EnvTab& tab) {
    Pixel<T> px;
    V3<T> nt;
    T A[3];
    load_pixel(px, nt, A, ai, vi.b, vi.pix, vi.plane);
    const size_t i3 = (vi.b * 3) * vi.plane + vi.pix;
    const size_t o3 = (vo.b * 3) * vo.plane + vo.pix;
    T gD[3], gS[3], gDa[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        gD[c] = (T)ai.g_diff[i3 + c * vi.plane];
        gS[c] = (T)ai.g_spec[i3 + c * vi.plane];
        gDa[c] = gD[c] * A[c] * Consts<T>::INV_PI;
    }
    PixelGrad<T> pg;
    pixel_grad_zero(pg);
    T gAlb[3] = {0, 0, 0};
    if (KP == KL) {
        // single sweep: all lobes in one chunk
        Lobes<T, KP> L;
        load_lobes(L, ai, vi.b, vi.pix, vi.plane, 0);
        T P[KP], Q[KP];
        load_PQ<T, KP>(P, Q, ai, vi.b, vi.pix, vi.plane, 0, gDa, gS);
        LobeGrads<T, KP> lg;
        sweep_bwd<T, KP, PIX, LOBE>(px, L, P, Q, tab, pg, lg, Coop::slice());
        if (PIX) coop_reduce<Coop>(pg);
        if (LOBE) coop_reduce<Coop>(lg);
        if (LOBE && Coop::leader()) store_lobe_grads<T, KP>(lg, L, P, Q, ai, vi, a, vo, 0, gDa, gS, gAlb);
    } else {
        if (PIX) {
            Lobes<T, KP> L;
            load_lobes(L, ai, vi.b, vi.pix, vi.plane, 0);
            T P[KP], Q[KP];
            load_PQ<T, KP>(P, Q, ai, vi.b, vi.pix, vi.plane, 0, gDa, gS);
            LobeGrads<T, KP> lg;  // unused
            sweep_bwd<T, KP, true, false>(px, L, P, Q, tab, pg, lg, Coop::slice());
            coop_reduce<Coop>(pg);
        }
        if (LOBE) {
            for (int k0 = 0; k0 < a.K; k0 += KL) {
                Lobes<T, KL> L;
                load_lobes(L, ai, vi.b, vi.pix, vi.plane, k0);
                T P[KL], Q[KL];
                load_PQ<T, KL>(P, Q, ai, vi.b, vi.pix, vi.plane, k0, gDa, gS);
                LobeGrads<T, KL> lg;
                sweep_bwd<T, KL, false, true>(px, L, P, Q, tab, pg, lg, Coop::slice());
                coop_reduce<Coop>(lg);
                if (Coop::leader()) store_lobe_grads<T, KL>(lg, L, P, Q, ai, vi, a, vo, k0, gDa, gS, gAlb);
            }
        }
    }
    if (!Coop::leader()) return;
    if (PIX) {
        V3<T> g_nt;
        T g_rough;
        pixel_bwd_finish(pg, px, nt, g_nt, g_rough);
        if (a.g_normal) {
            a.g_normal[o3] = (float)g_nt.x;
            a.g_normal[o3 + vo.plane] = (float)g_nt.y;
            a.g_normal[o3 + 2 * vo.plane] = (float)g_nt.z;
        }
        if (a.g_rough) a.g_rough[vo.b * vo.plane + vo.pix] = (float)g_rough;
    }
    if (LOBE && a.g_albedo) {
#pragma unroll
        for (int c = 0; c < 3; ++c) a.g_albedo[o3 + c * vo.plane] = (float)(gD[c] * Consts<T>::INV_PI * gAlb[c]);
    }
}
template <typename T, int KP, int KL, bool PIX, bool LOBE>
IRB_HD void drv_render_sg_bwd(const RenderSgArgs& a, const EnvTab& tab, size_t b, size_t pix) {
    const PixView v{b, pix, (size_t)a.R * a.C};
    drv_render_sg_bwd<T, KP, KL, PIX, LOBE>(a, v, a, v, tab);
}

}  // namespace irb
