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
template <int KC, bool PIX, bool LOBE>
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

template <typename T, int KC>
IRB_HD void sweep_fwd(const Pixel<T>& px, const Lobes<T, KC>& L, const EnvTab& tab, T (&U)[KC], T (&V)[KC]) {
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
            for (int d = 0; d < tab.D; ++d) {
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
        for (int d = 0; d < tab.D; ++d) {
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
        for (int d = 0; d < tab.D; ++d) {
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
                      PixelGrad<T>& pg, LobeGrads<T, KC>& lg) {
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
            for (int d = 0; d < tab.D; ++d) {
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
        for (int d = 0; d < tab.D; ++d) {
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
        for (int d = 0; d < tab.D; ++d) {
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
