// irb_pixel.cuh -- per-pixel direction sweeps of the fused SG render (forward and backward),
// register-resident over a chunk of KC lobes.  Shared by the CUDA kernels (float, device) and
// tests/hostcheck (float/double, host).
//
// Algebra (models.py:371-389 + 511-520), per pixel, with
//   e_kd = exp(lam_k (a_k.w_d - 1)),  w_d = Omega_d * ndl_d,  spec_d from dir_fwd*():
//   forward :  U_k = sum_d w_d e_kd,  V_k = sum_d w_d spec_d e_kd
//              diffuse_c = (A_c/pi) sum_k W_kc U_k,   spec_c = sum_k W_kc V_k
//   backward:  P_k = sum_c gD_c (A_c/pi) W_kc,  Q_k = sum_c gS_c W_kc
//              eD_d = Omega_d sum_k P_k e_kd,  eS_d = Omega_d sum_k Q_k e_kd   -> dir_bwd*()
//              T_kd = w_d e_kd (P_k + spec_d Q_k)
//              dL/dlam_k = sum_d T_kd (a_k.w_d - 1),   dL/da_k = lam_k sum_d T_kd w_d
//              dL/dW_kc = gD_c (A_c/pi) U_k + gS_c V_k,  dL/dA_c = (gD_c/pi) sum_k W_kc U_k
// i.e. the env-map E_cd is never formed: every quantity the BRDF side needs is linear in E.
//
// The exponent is evaluated with the sharpness folded into the axis,
//   arg_kd = fma(bx, wx, fma(by, wy, fma(bz, wz, -lam')))  with b = lam' a, lam' = lam log2(e),
// (3 FMA, no separate multiply) and  sum_d T (a.w - 1) = (sum_d T arg) / lam'.
#pragma once
#include "irb_math.cuh"
#if defined(__CUDACC__)
// the generic (host / double) loops behind the `if constexpr` device-float branches are dead code in float device
// instantiations by design
#pragma nv_diag_suppress 128
#endif

namespace irb {

#ifndef IRB_MAX_DIRS
#define IRB_MAX_DIRS 128
#endif

// hemisphere table: ls[d] (models.py:437-446) and envWeight[d] (models.py:448), fp32 values
struct EnvTab {
    float lx[IRB_MAX_DIRS], ly[IRB_MAX_DIRS], lz[IRB_MAX_DIRS], w[IRB_MAX_DIRS];
    int D;
};

template <typename T, int KC> struct Lobes {
    T bx[KC], by[KC], bz[KC], lam[KC];  // b = lam' * axis;  lam' = warped lam * log2(e)
};

template <typename T> IRB_HD T lobe_arg(T bx, T by, T bz, T lam, T wx, T wy, T wz) {
    return bx * wx + (by * wy + (bz * wz - lam));
}

// All KC exponentials of one direction, written in three explicit phases (arguments, SFU, uses) so that the
// KC independent chains are in flight together instead of one fma->fma->fma->ex2->use chain after the other
// (ncu: the dependent multiply right behind each MUFU.EX2 was the top short-scoreboard stall).
template <typename T, int KC>
IRB_HD void lobe_exps(const Lobes<T, KC>& L, T wx, T wy, T wz, T (&arg)[KC], T (&e)[KC]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) arg[k] = lobe_arg(L.bx[k], L.by[k], L.bz[k], L.lam[k], wx, wy, wz);
#pragma unroll
    for (int k = 0; k < KC; ++k) e[k] = x_exp2(arg[k]);
}

template <typename T, int KC> struct LobeGrads {
    T U[KC], V[KC], S[KC], mx[KC], my[KC], mz[KC];  // S = sum T arg ; m = sum T w_d
};

#if defined(__CUDACC__)
// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two lobes per instruction) ----------------
// The sweeps are issue-slot / FMA-pipe co-limited (ncu: 62 % issue, 50 % FMA pipe); one f32x2 instruction does
// the work of two scalar ones for one issue slot (profiles/microbench: FFMA2 sustains the full 128 lanes/clk/SM
// at half the issue rate).  Lobes are processed in pairs (k, k+1) held as 64-bit register pairs.
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk(float lo, float hi) {
    pk2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(pk2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pk2 fma2(pk2 a, pk2 b, pk2 c) {
    pk2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ pk2 mul2(pk2 a, pk2 b) {
    pk2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pk2 add2(pk2 a, pk2 b) {
    pk2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pk2 ex2x2(pk2 a) {
    float lo, hi;
    upk(a, lo, hi);
    return pk(x_exp2(lo), x_exp2(hi));
}

template <int KC> struct LobesPk {
    pk2 bx[KC / 2], by[KC / 2], bz[KC / 2], nlam[KC / 2];
};
template <int KC> __device__ __forceinline__ void pack_lobes(LobesPk<KC>& Lp, const Lobes<float, KC>& L) {
    static_assert(KC % 2 == 0, "lobe chunk must be even");
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) {
        Lp.bx[i] = pk(L.bx[2 * i], L.bx[2 * i + 1]);
        Lp.by[i] = pk(L.by[2 * i], L.by[2 * i + 1]);
        Lp.bz[i] = pk(L.bz[2 * i], L.bz[2 * i + 1]);
        Lp.nlam[i] = pk(-L.lam[2 * i], -L.lam[2 * i + 1]);
    }
}
// arguments and exponentials of all lobe pairs for one direction
template <int KC>
__device__ __forceinline__ void lobe_exps_pk(const LobesPk<KC>& Lp, pk2 wx2, pk2 wy2, pk2 wz2, pk2 (&arg)[KC / 2], pk2 (&e)[KC / 2]) {
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) arg[i] = fma2(Lp.bx[i], wx2, fma2(Lp.by[i], wy2, fma2(Lp.bz[i], wz2, Lp.nlam[i])));
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) e[i] = ex2x2(arg[i]);
}
template <int KC> struct LobeGradsPk {
    pk2 U[KC / 2], V[KC / 2], S[KC / 2], mx[KC / 2], my[KC / 2], mz[KC / 2];
};
template <int KC, bool PIX, bool LOBE>
__device__ __forceinline__ float lobes_bwd_dir_pk(LobeGradsPk<KC>& g, const LobesPk<KC>& Lp, const pk2 (&P2)[KC / 2],
                                                  const pk2 (&Q2)[KC / 2], float wx, float wy, float wz, float w, float spec) {
    const pk2 wx2 = pk(wx, wx), wy2 = pk(wy, wy), wz2 = pk(wz, wz);
    pk2 arg[KC / 2], e[KC / 2];
    lobe_exps_pk<KC>(Lp, wx2, wy2, wz2, arg, e);
    const pk2 w2 = pk(w, w), sp2 = pk(spec, spec);
    pk2 sQ2 = pk(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) {
        if (PIX) sQ2 = fma2(Q2[i], e[i], sQ2);
        if (LOBE) {
            const pk2 ew = mul2(e[i], w2);
            g.U[i] = add2(g.U[i], ew);
            g.V[i] = fma2(ew, sp2, g.V[i]);
            const pk2 t = mul2(ew, fma2(sp2, Q2[i], P2[i]));
            g.S[i] = fma2(t, arg[i], g.S[i]);
            g.mx[i] = fma2(t, wx2, g.mx[i]);
            g.my[i] = fma2(t, wy2, g.my[i]);
            g.mz[i] = fma2(t, wz2, g.mz[i]);
        }
    }
    float a, b;
    upk(sQ2, a, b);
    return a + b;
}
// accumulation phase only (arguments / exponentials precomputed by lobe_exps_pk)
template <int KC, bool PIX, bool LOBE, bool SAVED = false>
__device__ __forceinline__ float lobes_bwd_acc_pk(LobeGradsPk<KC>& g, const pk2 (&P2)[KC / 2], const pk2 (&Q2)[KC / 2],
                                                  const pk2 (&arg)[KC / 2], const pk2 (&e)[KC / 2], float wx, float wy, float wz,
                                                  float w, float spec) {
    const pk2 wx2 = pk(wx, wx), wy2 = pk(wy, wy), wz2 = pk(wz, wz);
    const pk2 w2 = pk(w, w), sp2 = pk(spec, spec);
    pk2 sQ2 = pk(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) {
        if (PIX) sQ2 = fma2(Q2[i], e[i], sQ2);
        if (LOBE) {
            const pk2 ew = mul2(e[i], w2);
            if (!SAVED) {   // SAVED: U_k, V_k come from the forward pass (RenderSgArgs::lobe_uv_in)
                g.U[i] = add2(g.U[i], ew);
                g.V[i] = fma2(ew, sp2, g.V[i]);
            }
            const pk2 t = mul2(ew, fma2(sp2, Q2[i], P2[i]));
            g.S[i] = fma2(t, arg[i], g.S[i]);
            g.mx[i] = fma2(t, wx2, g.mx[i]);
            g.my[i] = fma2(t, wy2, g.my[i]);
            g.mz[i] = fma2(t, wz2, g.mz[i]);
        }
    }
    float a, b;
    upk(sQ2, a, b);
    return a + b;
}
template <int KC> __device__ __forceinline__ void gradspk_zero(LobeGradsPk<KC>& g) {
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) g.U[i] = g.V[i] = g.S[i] = g.mx[i] = g.my[i] = g.mz[i] = pk(0.f, 0.f);
}
template <int KC> __device__ __forceinline__ void gradspk_unpack(LobeGrads<float, KC>& lg, const LobeGradsPk<KC>& g) {
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) {
        upk(g.U[i], lg.U[2 * i], lg.U[2 * i + 1]);
        upk(g.V[i], lg.V[2 * i], lg.V[2 * i + 1]);
        upk(g.S[i], lg.S[2 * i], lg.S[2 * i + 1]);
        upk(g.mx[i], lg.mx[2 * i], lg.mx[2 * i + 1]);
        upk(g.my[i], lg.my[2 * i], lg.my[2 * i + 1]);
        upk(g.mz[i], lg.mz[2 * i], lg.mz[2 * i + 1]);
    }
}
template <int KC> __device__ __forceinline__ void pack_pq(pk2 (&P2)[KC / 2], pk2 (&Q2)[KC / 2], const float (&P)[KC], const float (&Q)[KC]) {
#pragma unroll
    for (int i = 0; i < KC / 2; ++i) {
        P2[i] = pk(P[2 * i], P[2 * i + 1]);
        Q2[i] = pk(Q[2 * i], Q[2 * i + 1]);
    }
}
#endif  // __CUDACC__

#if defined(__CUDACC__)
// ---- packed (two directions per instruction) local-frame BRDF: dir_fwd_local / dir_bwd_local on f32x2 ----------
// Same formulas as irb_math.cuh, evaluated for directions (d0, d1) in the two halves of 64-bit register pairs; the
// special-function ops and the clamps/selects work on the halves.
struct DirL2 {
    pk2 Mx, My, Mz, rinv, vdh, p2, F, ndh_raw, ndh, n0, n2, nom_raw, inv_nom, spec;
    bool ok0, ok1;  // m.m clamp inactive (per half)
};
__device__ __forceinline__ pk2 pkb(float a) { return pk(a, a); }
__device__ __forceinline__ pk2 sel2(bool c0, bool c1, pk2 a, pk2 b) {
    float a0, a1, b0, b1;
    upk(a, a0, a1);
    upk(b, b0, b1);
    return pk(c0 ? a0 : b0, c1 ? a1 : b1);
}
__device__ __forceinline__ void dir_fwd_local2(DirL2& dr, const Pixel<float>& px, const PixelLocal<float>& pl, pk2 wx, pk2 wy, pk2 wz) {
    const pk2 vx = pkb(pl.vx), vy = pkb(pl.vy), vz = pkb(pl.vz);
    dr.Mx = add2(vx, wx);
    dr.My = add2(vy, wy);
    dr.Mz = add2(vz, wz);
    const pk2 q = fma2(dr.My, dr.My, mul2(dr.Mx, dr.Mx));
    const pk2 MM = fma2(dr.Mz, dr.Mz, q);
    float m0, m1;
    upk(MM, m0, m1);
    dr.ok0 = m0 >= 4e-6f;
    dr.ok1 = m1 >= 4e-6f;
    dr.rinv = pk(x_rsqrt(fmaxf(m0, 4e-6f)), x_rsqrt(fmaxf(m1, 4e-6f)));
    dr.vdh = mul2(fma2(vz, dr.Mz, fma2(vy, dr.My, mul2(vx, dr.Mx))), dr.rinv);
    dr.p2 = ex2x2(mul2(fma2(pkb(-5.55472f), dr.vdh, pkb(-6.98316f)), dr.vdh));
    dr.F = fma2(pkb(1.0f - px.F0), dr.p2, pkb(px.F0));
    dr.ndh_raw = mul2(dr.Mz, dr.rinv);
    float r0, r1;
    upk(dr.ndh_raw, r0, r1);
    dr.ndh = pk(__saturatef(r0), __saturatef(r1));
    const pk2 nn = mul2(dr.ndh, dr.ndh);
    const pk2 a2n = mul2(pkb(px.alpha2), nn);
    pk2 n0 = fma2(q, mul2(dr.rinv, dr.rinv), a2n);             // 1 - ndh^2 = (Mx^2+My^2)/|M|^2
    if (!(dr.ok0 && dr.ok1)) n0 = sel2(dr.ok0, dr.ok1, n0, add2(fma2(nn, pkb(-1.0f), a2n), pkb(1.0f)));
    float n00, n01;
    upk(n0, n00, n01);
    dr.n0 = pk(r0 <= 0.f ? 1.0f : (r0 >= 1.0f ? px.alpha2 : n00), r1 <= 0.f ? 1.0f : (r1 >= 1.0f ? px.alpha2 : n01));
    dr.n2 = fma2(wz, pkb(1.0f - px.k), pkb(px.k));
    dr.nom_raw = mul2(mul2(pkb(Consts<float>::FOUR_PI * px.n1), mul2(dr.n0, dr.n0)), dr.n2);
    float q0, q1;
    upk(dr.nom_raw, q0, q1);
    dr.inv_nom = pk(x_rcp(fminf(fmaxf(q0, 1e-6f), Consts<float>::FOUR_PI)), x_rcp(fminf(fmaxf(q1, 1e-6f), Consts<float>::FOUR_PI)));
    dr.spec = mul2(mul2(pkb(px.alpha2), dr.F), dr.inv_nom);
}

struct PixelGradL2 {
    pk2 Gxx, Gxy, Gxz, Gyx, Gyy, Gyz, Gzx, Gzy, Gzz, gnx, gny, gnz, G_n1, g_alpha2, g_k;
};
__device__ __forceinline__ void pixel_gradl2_zero(PixelGradL2& g) {
    const pk2 z = pk(0.f, 0.f);
    g.Gxx = g.Gxy = g.Gxz = g.Gyx = g.Gyy = g.Gyz = g.Gzx = g.Gzy = g.Gzz = z;
    g.gnx = g.gny = g.gnz = g.G_n1 = g.g_alpha2 = g.g_k = z;
}
__device__ __forceinline__ float hsum2(pk2 v) {
    float a, b;
    upk(v, a, b);
    return a + b;
}
__device__ __forceinline__ void pixel_gradl2_fold(PixelGradL<float>& o, const PixelGradL2& g) {
    o.Gxx = hsum2(g.Gxx); o.Gxy = hsum2(g.Gxy); o.Gxz = hsum2(g.Gxz);
    o.Gyx = hsum2(g.Gyx); o.Gyy = hsum2(g.Gyy); o.Gyz = hsum2(g.Gyz);
    o.Gzx = hsum2(g.Gzx); o.Gzy = hsum2(g.Gzy); o.Gzz = hsum2(g.Gzz);
    o.gnx = hsum2(g.gnx); o.gny = hsum2(g.gny); o.gnz = hsum2(g.gnz);
    o.G_n1 = hsum2(g.G_n1); o.g_alpha2 = hsum2(g.g_alpha2); o.g_k = hsum2(g.g_k);
}
// g_spec (two directions) = Omega_d ndl_d sum_k Q_k e_kd
__device__ __forceinline__ void dir_bwd_local2(PixelGradL2& g, const Pixel<float>& px, const PixelLocal<float>& pl, const DirL2& dr,
                                               pk2 wx, pk2 wy, pk2 wz, pk2 g_spec) {
    const pk2 gsi = mul2(g_spec, dr.inv_nom);
    const pk2 g_F = mul2(gsi, pkb(px.alpha2));
    g.g_alpha2 = fma2(gsi, dr.F, g.g_alpha2);
    float q0, q1;
    upk(dr.nom_raw, q0, q1);
    const bool p0 = (q0 >= 1e-6f) && (q0 <= Consts<float>::FOUR_PI), p1 = (q1 >= 1e-6f) && (q1 <= Consts<float>::FOUR_PI);
    pk2 c4 = mul2(mul2(gsi, dr.spec), pkb(-Consts<float>::FOUR_PI));   // g_nom * 4 pi
    if (!(p0 && p1)) c4 = sel2(p0, p1, c4, pk(0.f, 0.f));
    const pk2 x = mul2(c4, mul2(dr.n0, dr.n0));
    g.G_n1 = fma2(x, dr.n2, g.G_n1);
    g.g_k = fma2(mul2(x, pkb(px.n1)), fma2(wz, pkb(-1.0f), pkb(1.0f)), g.g_k);
    const pk2 g_n0 = mul2(mul2(c4, pkb(2.0f * px.n1)), mul2(dr.n0, dr.n2));
    const pk2 g_ndh = mul2(mul2(g_n0, pkb(2.0f * (px.alpha2 - 1.0f))), dr.ndh);
    g.g_alpha2 = fma2(g_n0, mul2(dr.ndh, dr.ndh), g.g_alpha2);
    const pk2 g_vdh = mul2(mul2(mul2(g_F, pkb((1.0f - px.F0) * Consts<float>::LN2)), dr.p2),
                           fma2(pkb(-2.0f * 5.55472f), dr.vdh, pkb(-6.98316f)));
    float r0, r1;
    upk(dr.ndh_raw, r0, r1);
    const bool h0 = (r0 >= 0.f) && (r0 <= 1.f), h1 = (r1 >= 0.f) && (r1 <= 1.f);
    pk2 g_ndhr = g_ndh;
    if (!(h0 && h1)) g_ndhr = sel2(h0, h1, g_ndh, pk(0.f, 0.f));
    const pk2 ghx = mul2(g_vdh, pkb(pl.vx)), ghy = mul2(g_vdh, pkb(pl.vy)), ghz = fma2(g_vdh, pkb(pl.vz), g_ndhr);
    const pk2 hx = mul2(dr.Mx, dr.rinv), hy = mul2(dr.My, dr.rinv), hz = mul2(dr.Mz, dr.rinv);
    g.gnx = fma2(g_ndhr, hx, g.gnx);
    g.gny = fma2(g_ndhr, hy, g.gny);
    g.gnz = fma2(g_ndhr, hz, g.gnz);
    pk2 hg = fma2(g_ndhr, dr.ndh_raw, mul2(g_vdh, dr.vdh));
    if (!(dr.ok0 && dr.ok1)) hg = sel2(dr.ok0, dr.ok1, hg, pk(0.f, 0.f));
    const pk2 nhg = mul2(hg, pkb(-1.0f));
    const pk2 glx = mul2(fma2(hx, nhg, ghx), dr.rinv), gly = mul2(fma2(hy, nhg, ghy), dr.rinv), glz = mul2(fma2(hz, nhg, ghz), dr.rinv);
    g.Gxx = fma2(wx, glx, g.Gxx); g.Gxy = fma2(wx, gly, g.Gxy); g.Gxz = fma2(wx, glz, g.Gxz);
    g.Gyx = fma2(wy, glx, g.Gyx); g.Gyy = fma2(wy, gly, g.Gyy); g.Gyz = fma2(wy, glz, g.Gyz);
    g.Gzx = fma2(wz, glx, g.Gzx); g.Gzy = fma2(wz, gly, g.Gzy); g.Gzz = fma2(wz, glz, g.Gzz);
}
#endif  // __CUDACC__

// The sweeps walk the directions d0, d0 + dstep, ... : (0, 1) for a thread that owns the pixel, (lane, 32) when a warp
// shares one pixel and adds the partial sums up afterwards (every accumulator is a plain sum over directions).
struct DirSlice {
    int d0, dstep;
};

template <typename T, int KC>
IRB_HD void sweep_fwd(const Pixel<T>& px, const Lobes<T, KC>& L, const EnvTab& tab, T (&U)[KC], T (&V)[KC],
                      DirSlice sl = DirSlice{0, 1}) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        U[k] = T(0);
        V[k] = T(0);
    }
    if (px.regular) {
        PixelLocal<T> pl;
        pixel_local_setup(pl, px);
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(T) == 4) {
            LobesPk<KC> Lp;
            pack_lobes<KC>(Lp, L);
            pk2 U2[KC / 2], V2[KC / 2];
#pragma unroll
            for (int i = 0; i < KC / 2; ++i) U2[i] = V2[i] = pk(0.f, 0.f);
#pragma unroll 2
            for (int d = sl.d0; d < tab.D; d += sl.dstep) {
                const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                DirL<float> dr;
                dir_fwd_local(dr, px, pl, wx, wy, wz);
                const float w = tab.w[d] * wz;
                const float ws = w * dr.spec;
                pk2 arg[KC / 2], e[KC / 2];
                lobe_exps_pk<KC>(Lp, pk(wx, wx), pk(wy, wy), pk(wz, wz), arg, e);
                const pk2 w2 = pk(w, w), ws2 = pk(ws, ws);
#pragma unroll
                for (int i = 0; i < KC / 2; ++i) {
                    U2[i] = fma2(w2, e[i], U2[i]);
                    V2[i] = fma2(ws2, e[i], V2[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < KC / 2; ++i) {
                upk(U2[i], U[2 * i], U[2 * i + 1]);
                upk(V2[i], V[2 * i], V[2 * i + 1]);
            }
            return;
        }
#endif
#pragma unroll 2
        for (int d = sl.d0; d < tab.D; d += sl.dstep) {
            const T wx = (T)tab.lx[d], wy = (T)tab.ly[d], wz = (T)tab.lz[d];
            DirL<T> dr;
            dir_fwd_local(dr, px, pl, wx, wy, wz);
            const T w = (T)tab.w[d] * wz;
            const T ws = w * dr.spec;
            T arg[KC], e[KC];
            lobe_exps<T, KC>(L, wx, wy, wz, arg, e);
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                U[k] += w * e[k];
                V[k] += ws * e[k];
            }
        }
    } else {
        for (int d = sl.d0; d < tab.D; d += sl.dstep) {
            const T wx = (T)tab.lx[d], wy = (T)tab.ly[d], wz = (T)tab.lz[d];
            Dir<T> dr;
            dir_fwd(dr, px, wx, wy, wz);
            const T w = (T)tab.w[d] * dr.ndl;
            const T ws = w * dr.spec;
            T arg[KC], e[KC];
            lobe_exps<T, KC>(L, wx, wy, wz, arg, e);
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                U[k] += w * e[k];
                V[k] += ws * e[k];
            }
        }
    }
}

// lobe part of one direction of the backward sweep; returns sum_k Q_k e_kd
template <typename T, int KC, bool PIX, bool LOBE>
IRB_HD T lobes_bwd_dir(LobeGrads<T, KC>& lg, const Lobes<T, KC>& L, const T (&P)[KC], const T (&Q)[KC], T wx, T wy, T wz, T w,
                       T spec, T& sP) {
    T arg[KC], e[KC];
    lobe_exps<T, KC>(L, wx, wy, wz, arg, e);
    T sQ = T(0);
    sP = T(0);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        if (PIX) {
            sQ += Q[k] * e[k];
            sP += P[k] * e[k];
        }
        if (LOBE) {
            const T ew = e[k] * w;
            lg.U[k] += ew;
            lg.V[k] += ew * spec;
            const T t = ew * (P[k] + spec * Q[k]);
            lg.S[k] += t * arg[k];
            lg.mx[k] += t * wx;
            lg.my[k] += t * wy;
            lg.mz[k] += t * wz;
        }
    }
    return sQ;
}

// One backward sweep.  PIX: accumulate the BRDF-side gradients (needs ALL lobes of the pixel in
// this chunk); LOBE: accumulate the SG-side gradients of this chunk.
template <typename T, int KC, bool PIX, bool LOBE>
IRB_HD void sweep_bwd(const Pixel<T>& px, const Lobes<T, KC>& L, const T (&P)[KC], const T (&Q)[KC], const EnvTab& tab,
                      PixelGrad<T>& pg, LobeGrads<T, KC>& lg, DirSlice sl = DirSlice{0, 1}) {
    if (LOBE) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            lg.U[k] = lg.V[k] = lg.S[k] = lg.mx[k] = lg.my[k] = lg.mz[k] = T(0);
        }
    }
    if (px.regular) {
        PixelLocal<T> pl;
        pixel_local_setup(pl, px);
        PixelGradL<T> gl;
        pixel_gradl_zero(gl);
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(T) == 4) {
            LobesPk<KC> Lp;
            pack_lobes<KC>(Lp, L);
            pk2 P2[KC / 2], Q2[KC / 2];
            pack_pq<KC>(P2, Q2, P, Q);
            LobeGradsPk<KC> g2;
            gradspk_zero<KC>(g2);
            for (int d = sl.d0; d < tab.D; d += sl.dstep) {
                const float wx = tab.lx[d], wy = tab.ly[d], wz = tab.lz[d];
                DirL<float> dr;
                dir_fwd_local(dr, px, pl, wx, wy, wz);
                const float w = tab.w[d] * wz;
                const float sQ = lobes_bwd_dir_pk<KC, PIX, LOBE>(g2, Lp, P2, Q2, wx, wy, wz, w, dr.spec);
                if (PIX) dir_bwd_local(gl, px, pl, dr, wx, wy, wz, w * sQ);
            }
            if (LOBE) gradspk_unpack<KC>(lg, g2);
            if (PIX) gradl_to_world(pg, gl, px);
            return;
        }
#endif
        for (int d = sl.d0; d < tab.D; d += sl.dstep) {
            const T wx = (T)tab.lx[d], wy = (T)tab.ly[d], wz = (T)tab.lz[d];
            DirL<T> dr;
            dir_fwd_local(dr, px, pl, wx, wy, wz);
            const T w = (T)tab.w[d] * wz;
            T sP;
            const T sQ = lobes_bwd_dir<T, KC, PIX, LOBE>(lg, L, P, Q, wx, wy, wz, w, dr.spec, sP);
            if (PIX) dir_bwd_local(gl, px, pl, dr, wx, wy, wz, w * sQ);
        }
        if (PIX) gradl_to_world(pg, gl, px);
    } else {
        for (int d = sl.d0; d < tab.D; d += sl.dstep) {
            const T wx = (T)tab.lx[d], wy = (T)tab.ly[d], wz = (T)tab.lz[d];
            Dir<T> dr;
            dir_fwd(dr, px, wx, wy, wz);
            const T om = (T)tab.w[d];
            const T w = om * dr.ndl;
            T eD;
            const T eS = lobes_bwd_dir<T, KC, PIX, LOBE>(lg, L, P, Q, wx, wy, wz, w, dr.spec, eD);
            if (PIX) dir_bwd(pg, px, dr, wx, wy, wz, om * eD, om * eS);
        }
    }
}

}  // namespace irb
