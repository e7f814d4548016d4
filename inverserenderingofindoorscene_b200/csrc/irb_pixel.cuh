// irb_pixel.cuh -- per-pixel direction sweeps of the fused SG render (forward and backward),
// register-resident over a chunk of KC lobes.  Shared by the CUDA kernels (float, device) and
// tests/hostcheck (float/double, host).
//
// Algebra (models.py:371-389 + 511-520), per pixel, with
//   e_kd = exp(lam_k (a_k.w_d - 1)),  w_d = Omega_d * ndl_d,  spec_d from dir_fwd():
//   forward :  U_k = sum_d w_d e_kd,  V_k = sum_d w_d spec_d e_kd
//              diffuse_c = (A_c/pi) sum_k W_kc U_k,   spec_c = sum_k W_kc V_k
//   backward:  P_k = sum_c gD_c (A_c/pi) W_kc,  Q_k = sum_c gS_c W_kc
//              eD_d = Omega_d sum_k P_k e_kd,  eS_d = Omega_d sum_k Q_k e_kd   -> dir_bwd()
//              T_kd = w_d e_kd (P_k + spec_d Q_k)
//              dL/dlam_k = sum_d T_kd (a_k.w_d - 1),   dL/da_k = lam_k sum_d T_kd w_d
//              dL/dW_kc = gD_c (A_c/pi) U_k + gS_c V_k,  dL/dA_c = (gD_c/pi) sum_k W_kc U_k
// i.e. the env-map E_cd is never formed: every quantity the BRDF side needs is linear in E.
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
    T ax[KC], ay[KC], az[KC], lam[KC];  // lam already warped and multiplied by log2(e)
};

// exp(lam (a.w - 1)) = exp2(lam' * dm1), lam' = lam*log2e
template <typename T> IRB_HD T lobe_dm1(T ax, T ay, T az, T wx, T wy, T wz) {
    return ax * wx + (ay * wy + (az * wz - T(1)));
}

template <typename T, int KC>
IRB_HD void sweep_fwd(const Pixel<T>& px, const Lobes<T, KC>& L, const EnvTab& tab, T (&U)[KC], T (&V)[KC]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        U[k] = T(0);
        V[k] = T(0);
    }
    for (int d = 0; d < tab.D; ++d) {
        const T wx = (T)tab.lx[d], wy = (T)tab.ly[d], wz = (T)tab.lz[d];
        Dir<T> dr;
        dir_fwd(dr, px, wx, wy, wz);
        const T w = (T)tab.w[d] * dr.ndl;
        const T ws = w * dr.spec;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const T e = x_exp2(L.lam[k] * lobe_dm1(L.ax[k], L.ay[k], L.az[k], wx, wy, wz));
            U[k] += w * e;
            V[k] += ws * e;
        }
    }
}

template <typename T, int KC> struct LobeGrads {
    T U[KC], V[KC], gl[KC], mx[KC], my[KC], mz[KC];  // gl = sum T dm1 ; m = sum T w_d
};

// One backward sweep.  PIX: accumulate the BRDF-side gradients (needs ALL lobes of the pixel in
// this chunk); LOBE: accumulate the SG-side gradients of this chunk.
template <typename T, int KC, bool PIX, bool LOBE>
IRB_HD void sweep_bwd(const Pixel<T>& px, const Lobes<T, KC>& L, const T (&P)[KC], const T (&Q)[KC], const EnvTab& tab,
                      PixelGrad<T>& pg, LobeGrads<T, KC>& lg) {
    if (LOBE) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            lg.U[k] = lg.V[k] = lg.gl[k] = lg.mx[k] = lg.my[k] = lg.mz[k] = T(0);
        }
    }
    for (int d = 0; d < tab.D; ++d) {
        const T wx = (T)tab.lx[d], wy = (T)tab.ly[d], wz = (T)tab.lz[d];
        Dir<T> dr;
        dir_fwd(dr, px, wx, wy, wz);
        const T om = (T)tab.w[d];
        const T w = om * dr.ndl;
        T eD = T(0), eS = T(0);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const T dm1 = lobe_dm1(L.ax[k], L.ay[k], L.az[k], wx, wy, wz);
            const T e = x_exp2(L.lam[k] * dm1);
            if (PIX) {
                eD += P[k] * e;
                eS += Q[k] * e;
            }
            if (LOBE) {
                const T ew = e * w;
                lg.U[k] += ew;
                lg.V[k] += ew * dr.spec;
                const T t = ew * (P[k] + dr.spec * Q[k]);
                lg.gl[k] += t * dm1;
                lg.mx[k] += t * wx;
                lg.my[k] += t * wy;
                lg.mz[k] += t * wz;
            }
        }
        if (PIX) dir_bwd(pg, px, dr, wx, wy, wz, om * eD, om * eS);
    }
}

}  // namespace irb
