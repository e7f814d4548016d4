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

template <typename T, int KC> struct LobeGrads {
    T U[KC], V[KC], S[KC], mx[KC], my[KC], mz[KC];  // S = sum T arg ; m = sum T w_d
};

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
