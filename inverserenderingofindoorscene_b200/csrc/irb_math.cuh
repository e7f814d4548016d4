// irb_math.cuh -- per-pixel / per-direction shading math of the render layer, shared by the
// unfused (render_env) and fused (render_sg) kernels.  Templated on the scalar type so that
// tests/hostcheck can instantiate the SAME code on the host in float and double and compare
// it with the oracle without a GPU (test infrastructure; the product only uses the float
// device instantiation).
//
// Follows /root/reference/models.py:461-522 (renderingLayer.forwardEnv); the hand-derived
// backward follows autograd of those lines (clamp pass-through inclusive, F.normalize eps
// semantics), SURVEY.md Appendix A.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define IRB_HD __host__ __device__ __forceinline__
#else
#define IRB_HD inline
#endif

namespace irb {

template <typename T> struct Consts;
template <> struct Consts<float> {
    static constexpr float PI = 3.14159265358979323846f;
    static constexpr float INV_PI = 0.31830988618379067154f;
    static constexpr float FOUR_PI = 12.566370614359172f;
    static constexpr float LN2 = 0.69314718055994530942f;
    static constexpr float LOG2E = 1.44269504088896340736f;
};
template <> struct Consts<double> {
    static constexpr double PI = 3.14159265358979323846;
    static constexpr double INV_PI = 0.31830988618379067154;
    static constexpr double FOUR_PI = 12.566370614359172;
    static constexpr double LN2 = 0.69314718055994530942;
    static constexpr double LOG2E = 1.44269504088896340736;
};

// ---- scalar primitives (device fp32 uses the SFU approximations where accuracy allows) ----
IRB_HD float x_exp2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}
IRB_HD double x_exp2(double x) { return exp2(x); }
IRB_HD float x_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;  // arguments are clamped away from zero by the callers: no denormal guard needed
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}
IRB_HD double x_rsqrt(double x) { return 1.0 / sqrt(x); }
IRB_HD float x_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;  // <= 1 ulp, no IEEE slow path (the once-per-pixel code was dominated by div/sqrt subroutine calls; profiles/r01)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}
IRB_HD double x_sqrt(double x) { return sqrt(x); }
IRB_HD float x_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}
IRB_HD double x_rcp(double x) { return 1.0 / x; }
// a / b: reciprocal-multiply on the device (<= 2 ulp), exact division on the host check
template <typename T> IRB_HD T x_div(T a, T b) {
#if defined(__CUDA_ARCH__)
    return a * x_rcp(b);
#else
    return a / b;
#endif
}
IRB_HD float x_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
IRB_HD float x_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
IRB_HD double x_mul(double a, double b) { return a * b; }
IRB_HD double x_add(double a, double b) { return a + b; }
IRB_HD float x_fma(float a, float b, float c) { return fmaf(a, b, c); }
IRB_HD double x_fma(double a, double b, double c) { return fma(a, b, c); }
template <typename T> IRB_HD T x_min(T a, T b) { return a < b ? a : b; }
template <typename T> IRB_HD T x_max(T a, T b) { return a > b ? a : b; }
template <typename T> IRB_HD T x_clamp(T x, T lo, T hi) { return x_min(x_max(x, lo), hi); }

template <typename T> struct V3 {
    T x, y, z;
};
template <typename T> IRB_HD T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> IRB_HD V3<T> cross(const V3<T>& a, const V3<T>& b) {
    return V3<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> IRB_HD V3<T> axpy(T a, const V3<T>& x, const V3<T>& y) {
    return V3<T>{a * x.x + y.x, a * x.y + y.y, a * x.z + y.z};
}
template <typename T> IRB_HD V3<T> scale(T a, const V3<T>& x) { return V3<T>{a * x.x, a * x.y, a * x.z}; }

// ---- tan warp of SG sharpness / intensity (models.py:396-400) ---------------------------
// y = tan(pi/2 * 0.999 * x).  Near x -> 1 the derivative reaches 4e5, so one ulp of the
// fp32 argument is ~4e-5 relative in y.  The complement t = 1 - 0.999 x is formed with 0.999
// split in two floats and y = cot(pi/2 * t), which keeps the result within ~3e-7 of the
// exact value for fp32 inputs (the fp64 oracle's value).
template <typename T> IRB_HD T sg_warp(T x);
template <> IRB_HD double sg_warp<double>(double x) { return tan(Consts<double>::PI / 2 * (0.999 * x)); }
// tan(z) on [0, pi/4]: z * Q(z^2), Q fitted for minimum relative error (1.3e-7 evaluated in fp32)
IRB_HD float tan_q(float z) {
    const float u = z * z;
    float q = 0.009265977889299393f;
    q = fmaf(q, u, 0.0033387888688594103f);
    q = fmaf(q, u, 0.02428046613931656f);
    q = fmaf(q, u, 0.053457971662282944f);
    q = fmaf(q, u, 0.13338157534599304f);
    q = fmaf(q, u, 0.3333318531513214f);
    q = fmaf(q, u, 1.0f);
    return z * q;
}
template <> IRB_HD float sg_warp<float>(float x) {
    const float c_hi = 0.999f;                                   // fl(0.999)
    const float c_lo = (float)(0.999 - (double)0.999f);          // residual
    const float HALF_PI = 1.57079632679489661923f;
    // lower half: tan(pi/2 * 0.999 x) directly; upper half: cot(pi/2 * (1 - 0.999 x)) -- both arguments
    // stay in [0, pi/4] where the condition number of tan is <= pi/2.
    const float a = fmaf(c_lo, x, c_hi * x);                     // 0.999 x
    float t = fmaf(-c_hi, x, 1.0f);
    t = fmaf(-c_lo, x, t);                                       // 1 - 0.999 x, single rounding of the exact value
    const bool low = a <= 0.5f;
    const float tz = tan_q(HALF_PI * (low ? a : t));
    return low ? tz : x_rcp(tz);   // rcp.approx: <= 1 ulp, no IEEE-division slow path
}
// d tan(c x)/dx = c (1 + tan^2)
template <typename T> IRB_HD T sg_warp_grad(T y) {
    return (T)(Consts<double>::PI / 2 * 0.999) * (T(1) + y * y);
}

// ---- per-pixel state ----------------------------------------------------------------------
template <typename T> struct Pixel {
    V3<T> n, camx, camy, v;   // shading normal, local frame, view vector
    T s, inv;                 // |n~|^2, 1/sqrt(clamp(s))
    T nn_dev;                 // |n|^2 - 1 in exact arithmetic (0 unless the s clamp is active)
    T un, pn, qn;             // up.n, clamped |p|, clamped |q|
    bool p_ok, q_ok;          // norms above F.normalize eps
    bool regular;             // unit normal, well-conditioned orthonormal frame: local-frame fast path is exact
    T rho, k, alpha2, ndv_raw, ndv, n1, F0;
};

// models.py:466-468, 477-480, 493-500
template <typename T>
IRB_HD void pixel_setup(Pixel<T>& px, const V3<T>& nt, T rough, const V3<T>& v, T F0) {
    px.v = v;
    px.F0 = F0;
    // (x^2 + y^2) + z^2 with separate roundings (no FMA contraction): the same value ATen's
    // torch.sum(n*n, dim=1) produces, so the clamp(s, 1e-6, 1) regime (models.py:467-468) is
    // decided exactly as the reference's fp32 run decides it.
    px.s = x_add(x_add(x_mul(nt.x, nt.x), x_mul(nt.y, nt.y)), x_mul(nt.z, nt.z));
    const T sc = x_clamp(px.s, (T)1e-6, (T)1);
    px.inv = x_rsqrt(sc);
    px.n = scale(px.inv, nt);
    if (px.s > T(1)) {
        // upper clamp: n = n~, |n|^2 - 1 = s - 1 with s summed without cancellation error
        const T xx = nt.x * nt.x, yy = nt.y * nt.y, zz = nt.z * nt.z;
        px.nn_dev = ((xx - T(1)) + yy + zz) + (x_fma(nt.x, nt.x, -xx) + x_fma(nt.y, nt.y, -yy) + x_fma(nt.z, nt.z, -zz));
    } else if (px.s < (T)1e-6) {
        px.nn_dev = px.s * (T)1e6 - T(1);
    } else {
        px.nn_dev = T(0);
    }
    px.un = px.n.y;
    const V3<T> p{-px.un * px.n.x, T(1) - px.un * px.n.y, -px.un * px.n.z};
    const T pl = x_sqrt(dot(p, p));
    px.p_ok = pl >= (T)1e-12;
    px.pn = x_max(pl, (T)1e-12);
    px.camy = scale(x_rcp(px.pn), p);
    const V3<T> q = cross(px.camy, px.n);
    const T ql = x_sqrt(dot(q, q));
    px.q_ok = ql >= (T)1e-12;
    px.qn = x_max(ql, (T)1e-12);
    px.camx = scale(-x_rcp(px.qn), q);
    px.rho = (rough + T(1)) * T(0.5);
    px.k = (px.rho + T(1)) * (px.rho + T(1)) * T(0.125);
    const T a = px.rho * px.rho;
    px.alpha2 = a * a;
    px.ndv_raw = dot(px.n, v);
    px.ndv = x_clamp(px.ndv_raw, T(0), T(1));
    px.n1 = px.ndv * (T(1) - px.k) + px.k;
    px.regular = (px.s >= (T)1e-6) && (px.s <= T(1)) && (pl >= (T)1e-2) && px.q_ok;
}

// per-direction forward quantities (kept for the backward of the same direction)
template <typename T> struct Dir {
    V3<T> l, h;
    T mm, rinv, vdh, p2, F, ndh_raw, ndh, ndl_raw, ndl, n0, n2, nom_raw, nom, spec;
};

// models.py:482-509.  (wx,wy,wz) = ls[d]
template <typename T> IRB_HD void dir_fwd(Dir<T>& dr, const Pixel<T>& px, T wx, T wy, T wz) {
    dr.l = V3<T>{wx * px.camx.x + wy * px.camy.x + wz * px.n.x, wx * px.camx.y + wy * px.camy.y + wz * px.n.y,
                 wx * px.camx.z + wy * px.camy.z + wz * px.n.z};
    const V3<T> m{(px.v.x + dr.l.x) * T(0.5), (px.v.y + dr.l.y) * T(0.5), (px.v.z + dr.l.z) * T(0.5)};
    dr.mm = dot(m, m);
    dr.rinv = x_rsqrt(x_max(dr.mm, (T)1e-6));
    dr.h = scale(dr.rinv, m);
    dr.vdh = dot(px.v, dr.h);
    dr.p2 = x_exp2((T(-5.55472) * dr.vdh - T(6.98316)) * dr.vdh);
    dr.F = px.F0 + (T(1) - px.F0) * dr.p2;
    dr.ndh_raw = dot(px.n, dr.h);
    dr.ndh = x_clamp(dr.ndh_raw, T(0), T(1));
    dr.ndl_raw = dot(px.n, dr.l);
    dr.ndl = x_clamp(dr.ndl_raw, T(0), T(1));
    // n0 = ndh^2 (alpha2-1) + 1 (models.py:504) cancels catastrophically for glossy pixels near the
    // highlight (n0 -> alpha2 ~ 1e-3).  Evaluate it as (1-ndh)(1+ndh) + alpha2 ndh^2 with
    // 1-ndh = |n-h|^2/2 + (1 - (|n|^2+|h|^2)/2), which has no cancellation; px.nn_dev and the mm
    // clamp carry the (rare) non-unit cases exactly.
    if (dr.ndh_raw <= T(0)) {
        dr.n0 = T(1);
    } else if (dr.ndh_raw >= T(1)) {
        dr.n0 = px.alpha2;
    } else {
        const T dx = px.n.x - dr.h.x, dy = px.n.y - dr.h.y, dz = px.n.z - dr.h.z;
        const T hh_dev = (dr.mm >= (T)1e-6) ? T(0) : dr.mm * (T)1e6 - T(1);
        const T one_m = T(0.5) * (dx * dx + dy * dy + dz * dz) - T(0.5) * (px.nn_dev + hh_dev);
        dr.n0 = one_m * (T(1) + dr.ndh) + px.alpha2 * dr.ndh * dr.ndh;
    }
    dr.n2 = dr.ndl * (T(1) - px.k) + px.k;
    dr.nom_raw = Consts<T>::FOUR_PI * dr.n0 * dr.n0 * px.n1 * dr.n2;
    dr.nom = x_clamp(dr.nom_raw, (T)1e-6, Consts<T>::FOUR_PI);
    dr.spec = px.alpha2 * dr.F * x_rcp(dr.nom);
}

// accumulators of the per-pixel backward
template <typename T> struct PixelGrad {
    V3<T> g_cx, g_cy, g_n;
    T G_n1, g_alpha2, g_k;
};
template <typename T> IRB_HD void pixel_grad_zero(PixelGrad<T>& g) {
    g.g_cx = V3<T>{0, 0, 0};
    g.g_cy = V3<T>{0, 0, 0};
    g.g_n = V3<T>{0, 0, 0};
    g.G_n1 = 0;
    g.g_alpha2 = 0;
    g.g_k = 0;
}

// Backward of one direction.  eD = Omega_d * sum_c gD_c (A_c/pi) E_cd, eS = Omega_d * sum_c gS_c E_cd.
template <typename T>
IRB_HD void dir_bwd(PixelGrad<T>& g, const Pixel<T>& px, const Dir<T>& dr, T wx, T wy, T wz, T eD, T eS) {
    T g_ndl = eD + eS * dr.spec;
    const T g_spec = eS * dr.ndl;
    const T inv_nom = x_rcp(dr.nom);
    const T g_F = g_spec * px.alpha2 * inv_nom;
    const bool nom_pass = (dr.nom_raw >= (T)1e-6) && (dr.nom_raw <= Consts<T>::FOUR_PI);
    const T g_nom = nom_pass ? -g_spec * dr.spec * inv_nom : T(0);
    g.g_alpha2 += g_spec * dr.F * inv_nom;
    const T c4 = g_nom * Consts<T>::FOUR_PI;
    const T g_n0 = c4 * T(2) * dr.n0 * px.n1 * dr.n2;
    g.G_n1 += c4 * dr.n0 * dr.n0 * dr.n2;
    const T g_n2 = c4 * dr.n0 * dr.n0 * px.n1;
    const T g_ndh = g_n0 * T(2) * dr.ndh * (px.alpha2 - T(1));
    g.g_alpha2 += g_n0 * dr.ndh * dr.ndh;
    g_ndl += g_n2 * (T(1) - px.k);
    g.g_k += g_n2 * (T(1) - dr.ndl);
    const T g_g = g_F * (T(1) - px.F0) * dr.p2 * Consts<T>::LN2;
    const T g_vdh = g_g * (T(-2 * 5.55472) * dr.vdh - T(6.98316));
    const T g_ndhr = (dr.ndh_raw >= T(0) && dr.ndh_raw <= T(1)) ? g_ndh : T(0);
    const V3<T> g_h{g_ndhr * px.n.x + g_vdh * px.v.x, g_ndhr * px.n.y + g_vdh * px.v.y,
                    g_ndhr * px.n.z + g_vdh * px.v.z};
    g.g_n = axpy(g_ndhr, dr.h, g.g_n);
    V3<T> g_m;
    if (dr.mm >= (T)1e-6) {
        const T hg = dot(dr.h, g_h);
        g_m = V3<T>{(g_h.x - dr.h.x * hg) * dr.rinv, (g_h.y - dr.h.y * hg) * dr.rinv, (g_h.z - dr.h.z * hg) * dr.rinv};
    } else {
        g_m = scale(dr.rinv, g_h);
    }
    const T g_ndlr = (dr.ndl_raw >= T(0) && dr.ndl_raw <= T(1)) ? g_ndl : T(0);
    const V3<T> g_l{T(0.5) * g_m.x + g_ndlr * px.n.x, T(0.5) * g_m.y + g_ndlr * px.n.y, T(0.5) * g_m.z + g_ndlr * px.n.z};
    g.g_n = axpy(g_ndlr, dr.l, g.g_n);
    g.g_cx = axpy(wx, g_l, g.g_cx);
    g.g_cy = axpy(wy, g_l, g.g_cy);
    g.g_n = axpy(wz, g_l, g.g_n);
}

// After the direction loop: frame / normalisation / roughness chain.  Returns d/d(pooled normal)
// and d/d(pooled rough).
template <typename T>
IRB_HD void pixel_bwd_finish(const PixelGrad<T>& g, const Pixel<T>& px, const V3<T>& nt, V3<T>& g_nt, T& g_rough) {
    V3<T> g_n = g.g_n;
    V3<T> g_cy = g.g_cy;
    const T g_ndv = g.G_n1 * (T(1) - px.k);
    const T g_k = g.g_k + g.G_n1 * (T(1) - px.ndv);
    if (px.ndv_raw >= T(0) && px.ndv_raw <= T(1)) g_n = axpy(g_ndv, px.v, g_n);
    // camx = -q / qn
    V3<T> g_q;
    if (px.q_ok) {
        const T cg = dot(px.camx, g.g_cx);
        const T iq = x_rcp(px.qn);
        g_q = V3<T>{-(g.g_cx.x - px.camx.x * cg) * iq, -(g.g_cx.y - px.camx.y * cg) * iq, -(g.g_cx.z - px.camx.z * cg) * iq};
    } else {
        g_q = scale(-x_rcp(px.qn), g.g_cx);
    }
    // q = camy x n
    const V3<T> t1 = cross(px.n, g_q);
    g_cy = V3<T>{g_cy.x + t1.x, g_cy.y + t1.y, g_cy.z + t1.z};
    const V3<T> t2 = cross(g_q, px.camy);
    g_n = V3<T>{g_n.x + t2.x, g_n.y + t2.y, g_n.z + t2.z};
    // camy = p / pn
    V3<T> g_p;
    if (px.p_ok) {
        const T cg = dot(px.camy, g_cy);
        const T ip = x_rcp(px.pn);
        g_p = V3<T>{(g_cy.x - px.camy.x * cg) * ip, (g_cy.y - px.camy.y * cg) * ip, (g_cy.z - px.camy.z * cg) * ip};
    } else {
        g_p = scale(x_rcp(px.pn), g_cy);
    }
    // p = up - (up.n) n
    const T ngp = dot(px.n, g_p);
    g_n = V3<T>{g_n.x - px.un * g_p.x, g_n.y - px.un * g_p.y - ngp, g_n.z - px.un * g_p.z};
    // n = n~ / sqrt(clamp(s))
    if (px.s >= (T)1e-6 && px.s <= T(1)) {
        const T ng = dot(px.n, g_n);
        g_nt = V3<T>{(g_n.x - px.n.x * ng) * px.inv, (g_n.y - px.n.y * ng) * px.inv, (g_n.z - px.n.z * ng) * px.inv};
    } else {
        g_nt = scale(px.inv, g_n);
    }
    (void)nt;
    const T g_rho = g.g_alpha2 * T(4) * px.rho * px.rho * px.rho + g_k * (px.rho + T(1)) * T(0.25);
    g_rough = g_rho * T(0.5);
}


// =============================================================================================
// Regular-regime fast path.  When 1e-6 <= |n~|^2 <= 1 (every real caller: normals are pooled
// means of unit vectors) the frame (camx, camy, n) is orthonormal, so the shading of
// models.py:482-509 can be evaluated in the pixel's LOCAL frame where n = (0,0,1), l = ls[d]
// and n.l = ls[d].z is a constant: no world-space l/h vectors, 1 - ndh^2 = (Mx^2+My^2)/|M|^2
// without cancellation, and d(n.l)/d(normal) = 0 so that whole gradient branch disappears
// (SURVEY.md 8a "derived identities").  ~22 FMA-pipe ops + 3 SFU ops per direction forward,
// ~45 backward, against ~70 / ~190 for the general path, which remains the fallback for
// clamped / degenerate pixels and is what the reference's autograd does there.
// =============================================================================================
template <typename T> struct PixelLocal {
    T vx, vy, vz;  // view vector in the (camx, camy, n) frame
};
template <typename T> IRB_HD void pixel_local_setup(PixelLocal<T>& pl, const Pixel<T>& px) {
    pl.vx = dot(px.v, px.camx);
    pl.vy = dot(px.v, px.camy);
    pl.vz = dot(px.v, px.n);
}

template <typename T> struct DirL {
    T Mx, My, Mz, rinv, vdh, p2, F, ndh_raw, ndh, n0, n2, nom_raw, inv_nom, spec;
    bool mm_ok;
};

template <typename T> IRB_HD void dir_fwd_local(DirL<T>& dr, const Pixel<T>& px, const PixelLocal<T>& pl, T wx, T wy, T wz) {
    dr.Mx = pl.vx + wx;
    dr.My = pl.vy + wy;
    dr.Mz = pl.vz + wz;                       // M = v + l = 2 m
    const T q = dr.Mx * dr.Mx + dr.My * dr.My;
    const T MM = q + dr.Mz * dr.Mz;           // 4 m.m
    dr.mm_ok = MM >= (T)4e-6;                 // clamp(m.m, min=1e-6), models.py:487
    dr.rinv = x_rsqrt(x_max(MM, (T)4e-6));    // h = M * rinv
    dr.vdh = (pl.vx * dr.Mx + pl.vy * dr.My + pl.vz * dr.Mz) * dr.rinv;
    dr.p2 = x_exp2((T(-5.55472) * dr.vdh - T(6.98316)) * dr.vdh);
    dr.F = px.F0 + (T(1) - px.F0) * dr.p2;
    dr.ndh_raw = dr.Mz * dr.rinv;
    dr.ndh = x_clamp(dr.ndh_raw, T(0), T(1));
    // branch-free: 1 - ndh^2 = (Mx^2+My^2)/|M|^2 (no cancellation); under the m.m clamp |h| < 1 and the
    // literal form is used; the ndh clamps give n0 = 1 / alpha2 exactly
    const T a2n = px.alpha2 * dr.ndh * dr.ndh;
    const T n0_reg = dr.mm_ok ? (q * dr.rinv * dr.rinv + a2n) : (a2n - dr.ndh * dr.ndh + T(1));
    dr.n0 = dr.ndh_raw <= T(0) ? T(1) : (dr.ndh_raw >= T(1) ? px.alpha2 : n0_reg);
    dr.n2 = wz * (T(1) - px.k) + px.k;        // ndl = l.z
    dr.nom_raw = Consts<T>::FOUR_PI * dr.n0 * dr.n0 * px.n1 * dr.n2;
    dr.inv_nom = x_rcp(x_clamp(dr.nom_raw, (T)1e-6, Consts<T>::FOUR_PI));
    dr.spec = px.alpha2 * dr.F * dr.inv_nom;
}

// local-frame gradient accumulators: G[a][b] = sum_d ls[d].a * (dL/dl_d).b in frame coordinates,
// gn = sum_d g_ndh h_d (frame coordinates)
template <typename T> struct PixelGradL {
    T Gxx, Gxy, Gxz, Gyx, Gyy, Gyz, Gzx, Gzy, Gzz, gnx, gny, gnz, G_n1, g_alpha2, g_k;
};
template <typename T> IRB_HD void pixel_gradl_zero(PixelGradL<T>& g) {
    g.Gxx = g.Gxy = g.Gxz = g.Gyx = g.Gyy = g.Gyz = g.Gzx = g.Gzy = g.Gzz = T(0);
    g.gnx = g.gny = g.gnz = g.G_n1 = g.g_alpha2 = g.g_k = T(0);
}

// g_spec = dL/dspec_d = Omega_d * ndl_d * sum_c gS_c E_cd
template <typename T>
IRB_HD void dir_bwd_local(PixelGradL<T>& g, const Pixel<T>& px, const PixelLocal<T>& pl, const DirL<T>& dr, T wx, T wy, T wz,
                          T g_spec) {
    const T gsi = g_spec * dr.inv_nom;
    const T g_F = gsi * px.alpha2;
    g.g_alpha2 += gsi * dr.F;
    const bool nom_pass = (dr.nom_raw >= (T)1e-6) && (dr.nom_raw <= Consts<T>::FOUR_PI);
    const T c4 = nom_pass ? -gsi * dr.spec * Consts<T>::FOUR_PI : T(0);   // g_nom * 4 pi
    const T x = c4 * dr.n0 * dr.n0;
    g.G_n1 += x * dr.n2;
    g.g_k += x * px.n1 * (T(1) - wz);
    const T g_n0 = c4 * T(2) * dr.n0 * px.n1 * dr.n2;
    const T g_ndh = g_n0 * T(2) * dr.ndh * (px.alpha2 - T(1));
    g.g_alpha2 += g_n0 * dr.ndh * dr.ndh;
    const T g_vdh = g_F * (T(1) - px.F0) * dr.p2 * Consts<T>::LN2 * (T(-2 * 5.55472) * dr.vdh - T(6.98316));
    const T g_ndhr = (dr.ndh_raw >= T(0) && dr.ndh_raw <= T(1)) ? g_ndh : T(0);
    // g_h = g_ndhr * n + g_vdh * v   (frame coordinates, n = (0,0,1))
    const T ghx = g_vdh * pl.vx, ghy = g_vdh * pl.vy, ghz = g_vdh * pl.vz + g_ndhr;
    const T hx = dr.Mx * dr.rinv, hy = dr.My * dr.rinv, hz = dr.Mz * dr.rinv;
    g.gnx += g_ndhr * hx;
    g.gny += g_ndhr * hy;
    g.gnz += g_ndhr * hz;
    // h = M * rinv, l = M - v:  dL/dl = (g_h - h (h.g_h)) * rinv   (or g_h * rinv under the m.m clamp)
    const T hg = dr.mm_ok ? (g_vdh * dr.vdh + g_ndhr * dr.ndh_raw) : T(0);
    const T glx = (ghx - hx * hg) * dr.rinv, gly = (ghy - hy * hg) * dr.rinv, glz = (ghz - hz * hg) * dr.rinv;
    g.Gxx += wx * glx; g.Gxy += wx * gly; g.Gxz += wx * glz;
    g.Gyx += wy * glx; g.Gyy += wy * gly; g.Gyz += wy * glz;
    g.Gzx += wz * glx; g.Gzy += wz * gly; g.Gzz += wz * glz;
}

// frame coordinates -> world-space accumulators of the general backward
template <typename T> IRB_HD void gradl_to_world(PixelGrad<T>& w, const PixelGradL<T>& g, const Pixel<T>& px) {
    const V3<T>&cx = px.camx, &cy = px.camy, &n = px.n;
    w.g_cx = V3<T>{g.Gxx * cx.x + g.Gxy * cy.x + g.Gxz * n.x, g.Gxx * cx.y + g.Gxy * cy.y + g.Gxz * n.y,
                   g.Gxx * cx.z + g.Gxy * cy.z + g.Gxz * n.z};
    w.g_cy = V3<T>{g.Gyx * cx.x + g.Gyy * cy.x + g.Gyz * n.x, g.Gyx * cx.y + g.Gyy * cy.y + g.Gyz * n.y,
                   g.Gyx * cx.z + g.Gyy * cy.z + g.Gyz * n.z};
    const T ax = g.Gzx + g.gnx, ay = g.Gzy + g.gny, az = g.Gzz + g.gnz;
    w.g_n = V3<T>{ax * cx.x + ay * cy.x + az * n.x, ax * cx.y + ay * cy.y + az * n.y, ax * cx.z + ay * cy.z + az * n.z};
    w.G_n1 = g.G_n1;
    w.g_alpha2 = g.g_alpha2;
    w.g_k = g.g_k;
}

}  // namespace irb
