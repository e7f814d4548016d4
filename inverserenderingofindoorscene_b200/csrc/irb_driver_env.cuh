// irb_driver_env.cuh -- per-pixel drivers of the UNFUSED boundary kernels (materialised
// env-map [B,3,R,C,envH,envW]): sg_to_env fwd/bwd (models.py:371-404) and render_env fwd/bwd
// (models.py:466-520).  Thread-per-pixel; the env-map is walked four directions at a time
// (16-byte vector accesses; a pixel's 128 directions of one channel are 512 contiguous bytes).
// Host-instantiable for tests/hostcheck.
#pragma once
#include "irb_driver.cuh"

namespace irb {

struct F4 {
    float x, y, z, w;
};
IRB_HD F4 ld4(const float* p) {
#if defined(__CUDA_ARCH__)
    const float4 v = *reinterpret_cast<const float4*>(p);
    return F4{v.x, v.y, v.z, v.w};
#else
    return F4{p[0], p[1], p[2], p[3]};
#endif
}
IRB_HD void st4(float* p, float a, float b, float c, float d) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
#else
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

struct SgEnvArgs {
    const float* axis;    // [B,K,3,R,C]
    const float* lamb;    // [B,K,R,C]
    const float* weight;  // [B,3K,R,C]
    float* env;           // [B,3,R,C,D]
    float* lamb_w;        // optional warped outputs
    float* weight_w;
    const float* g_env;
    float* g_axis;
    float* g_lamb;
    float* g_weight;
    int B, K, R, C;
    int warp;
    // fused reconstruction loss (irb_sg_reconst_*; wrapperBRDFLight.py:179-188)
    const float* env_gt;   // [B,3,R,C,D]
    const float* seg;      // [B,1,R,C] mask / weight of the pixel
    double* stats;         // [B,4]: sum(E G s^2), sum(E^2 s^2), sum((log(cE+off)-log(G+off))^2 s), sum(s)
    const float* g_num;    // device scalar: dL/d(numerator)
    float offset;
};

IRB_HD RenderSgArgs as_render_args(const SgEnvArgs& s) {
    RenderSgArgs a{};
    a.axis = s.axis;
    a.lamb = s.lamb;
    a.weight = s.weight;
    a.B = s.B;
    a.K = s.K;
    a.R = s.R;
    a.C = s.C;
    a.warp = s.warp;
    return a;
}

// ---- sg_to_env forward (models.py:371-389).  D must be a multiple of 4. -------------------
template <typename T, int KC> IRB_HD void drv_sg_to_env_fwd(const SgEnvArgs& s, const EnvTab& tab, size_t b, size_t pix) {
    const RenderSgArgs a = as_render_args(s);
    const size_t plane = (size_t)s.R * s.C;
    const int D = tab.D;
    for (int k0 = 0; k0 < s.K; k0 += KC) {
        Lobes<T, KC> L;
        load_lobes(L, a, b, pix, plane, k0);
        T W[KC][3];
#pragma unroll
        for (int k = 0; k < KC; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) W[k][c] = (k0 + k < s.K) ? load_weight<T>(a, b, pix, plane, k0 + k, c) : T(0);
        if (s.lamb_w || s.weight_w) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if (k0 + k < s.K) {
                    if (s.lamb_w) s.lamb_w[(b * s.K + k0 + k) * plane + pix] = (float)(L.lam[k] * Consts<T>::LN2);
                    if (s.weight_w)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            s.weight_w[(b * (3 * s.K) + 3 * (k0 + k) + c) * plane + pix] = (float)W[k][c];
                }
            }
        }
        for (int d0 = 0; d0 < D; d0 += 4) {
            T E[3][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const T wx = (T)tab.lx[d0 + u], wy = (T)tab.ly[d0 + u], wz = (T)tab.lz[d0 + u];
                E[0][u] = E[1][u] = E[2][u] = T(0);
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const T e = x_exp2(lobe_arg(L.bx[k], L.by[k], L.bz[k], L.lam[k], wx, wy, wz));
                    E[0][u] += W[k][0] * e;
                    E[1][u] += W[k][1] * e;
                    E[2][u] += W[k][2] * e;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float* p = s.env + ((b * 3 + c) * plane + pix) * D + d0;
                if (k0 > 0) {  // later lobe chunks accumulate onto the first
                    const F4 o = ld4(p);
                    st4(p, (float)(E[c][0] + (T)o.x), (float)(E[c][1] + (T)o.y), (float)(E[c][2] + (T)o.z),
                        (float)(E[c][3] + (T)o.w));
                } else {
                    st4(p, (float)E[c][0], (float)E[c][1], (float)E[c][2], (float)E[c][3]);
                }
            }
        }
    }
}

// ---- sg_to_env backward: dL/dW_kc = sum_d G_cd e_kd; T_kd = e_kd sum_c G_cd W_kc; ... ------
template <typename T, int KC> IRB_HD void drv_sg_to_env_bwd(const SgEnvArgs& s, const EnvTab& tab, size_t b, size_t pix) {
    const RenderSgArgs a = as_render_args(s);
    const size_t plane = (size_t)s.R * s.C;
    const int D = tab.D;
    for (int k0 = 0; k0 < s.K; k0 += KC) {
        Lobes<T, KC> L;
        load_lobes(L, a, b, pix, plane, k0);
        T W[KC][3], gW[KC][3], gl[KC], nn[KC], mx[KC], my[KC], mz[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            gl[k] = nn[k] = mx[k] = my[k] = mz[k] = T(0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                W[k][c] = (k0 + k < s.K) ? load_weight<T>(a, b, pix, plane, k0 + k, c) : T(0);
                gW[k][c] = T(0);
            }
        }
        for (int d0 = 0; d0 < D; d0 += 4) {
            const F4 g0 = ld4(s.g_env + ((b * 3 + 0) * plane + pix) * D + d0);
            const F4 g1 = ld4(s.g_env + ((b * 3 + 1) * plane + pix) * D + d0);
            const F4 g2 = ld4(s.g_env + ((b * 3 + 2) * plane + pix) * D + d0);
            const T G[3][4] = {{(T)g0.x, (T)g0.y, (T)g0.z, (T)g0.w},
                               {(T)g1.x, (T)g1.y, (T)g1.z, (T)g1.w},
                               {(T)g2.x, (T)g2.y, (T)g2.z, (T)g2.w}};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const T wx = (T)tab.lx[d0 + u], wy = (T)tab.ly[d0 + u], wz = (T)tab.lz[d0 + u];
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const T arg = lobe_arg(L.bx[k], L.by[k], L.bz[k], L.lam[k], wx, wy, wz);
                    const T e = x_exp2(arg);
                    gW[k][0] += G[0][u] * e;
                    gW[k][1] += G[1][u] * e;
                    gW[k][2] += G[2][u] * e;
                    const T t = e * (G[0][u] * W[k][0] + G[1][u] * W[k][1] + G[2][u] * W[k][2]);
                    gl[k] += t * arg;
                    nn[k] += t;
                    mx[k] += t * wx;
                    my[k] += t * wy;
                    mz[k] += t * wz;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int kk = k0 + k;
            if (kk < s.K) {
                const T lam = L.lam[k] * Consts<T>::LN2;
                const size_t oa = ((b * s.K + kk) * 3) * plane + pix;
                if (s.g_axis) {
                    s.g_axis[oa] = (float)(lam * mx[k]);
                    s.g_axis[oa + plane] = (float)(lam * my[k]);
                    s.g_axis[oa + 2 * plane] = (float)(lam * mz[k]);
                }
                if (s.g_lamb) {
                    T g = lobe_lam_grad<T>(gl[k], L.lam[k], mx[k], my[k], mz[k], nn[k], s.axis, oa, plane);
                    if (s.warp) g *= sg_warp_grad<T>(lam);
                    s.g_lamb[(b * s.K + kk) * plane + pix] = (float)g;
                }
                if (s.g_weight) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        T g = gW[k][c];
                        if (s.warp) g *= sg_warp_grad<T>(W[k][c]);
                        s.g_weight[(b * (3 * s.K) + 3 * kk + c) * plane + pix] = (float)g;
                    }
                }
            }
        }
    }
}

// ---- render_env (models.py:466-520) ---------------------------------------------------------
struct RenderEnvArgs {
    const float* albedo;  // [B,3,R,C] pooled
    const float* normal;
    const float* rough;
    const float* env;   // [B,3,R,C,D]
    const float* view;  // [3,R,C]
    float* out_diff;
    float* out_spec;
    const float* g_diff;
    const float* g_spec;
    float* g_albedo;
    float* g_normal;
    float* g_rough;
    float* g_env;
    int B, R, C;
    float F0;
};

IRB_HD RenderSgArgs as_render_args(const RenderEnvArgs& s) {
    RenderSgArgs a{};
    a.albedo = s.albedo;
    a.normal = s.normal;
    a.rough = s.rough;
    a.view = s.view;
    a.B = s.B;
    a.R = s.R;
    a.C = s.C;
    a.F0 = s.F0;
    return a;
}

// Per-pixel state of render_env: begin / one call per block of 4 directions / end.  Shared by the plain thread-per-pixel
// drivers below (global 16-byte loads) and the TMA-staged kernels of irb_render_env_tma.cuh (env tiles in shared memory).
// Regular pixels (irb_math.cuh: unit normal, well-conditioned frame -- every real caller) shade in the local frame like the
// fused kernels: n.l = w_d.z is a constant and the per-direction chain is half as long as the world-space one.
template <typename T> struct RenderEnvPixel {
    Pixel<T> px;
    V3<T> nt;
    T A[3];
    PixelLocal<T> pl;
    T Dc[3], Sc[3];                       // forward accumulators
    T gD[3], gS[3], gDa[3], gAlb[3];      // backward
    PixelGrad<T> pg;
    PixelGradL<T> gl;

    IRB_HD void begin(const RenderEnvArgs& s, size_t b, size_t pix, size_t plane, bool bwd) {
        const RenderSgArgs a = as_render_args(s);
        load_pixel(px, nt, A, a, b, pix, plane);
        if (px.regular) pixel_local_setup(pl, px);
#pragma unroll
        for (int c = 0; c < 3; ++c) Dc[c] = Sc[c] = gAlb[c] = T(0);
        if (bwd) {
            const size_t o3 = (b * 3) * plane + pix;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                gD[c] = (T)s.g_diff[o3 + c * plane];
                gS[c] = (T)s.g_spec[o3 + c * plane];
                gDa[c] = gD[c] * A[c] * Consts<T>::INV_PI;
            }
            pixel_grad_zero(pg);
            pixel_gradl_zero(gl);
        }
    }
    // E[c][u] = env[c][d0 + u]
    IRB_HD void quad_fwd(const EnvTab& tab, int d0, const T (&E)[3][4]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            T w, ws;
            if (px.regular) {
                DirL<T> dl;
                dir_fwd_local(dl, px, pl, (T)tab.lx[d0 + u], (T)tab.ly[d0 + u], (T)tab.lz[d0 + u]);
                w = (T)tab.w[d0 + u] * (T)tab.lz[d0 + u];
                ws = w * dl.spec;
            } else {
                Dir<T> dr;
                dir_fwd(dr, px, (T)tab.lx[d0 + u], (T)tab.ly[d0 + u], (T)tab.lz[d0 + u]);
                w = (T)tab.w[d0 + u] * dr.ndl;
                ws = w * dr.spec;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Dc[c] += w * E[c][u];
                Sc[c] += ws * E[c][u];
            }
        }
    }
    IRB_HD void end_fwd(const RenderEnvArgs& s, size_t b, size_t pix, size_t plane) const {
        const size_t o3 = (b * 3) * plane + pix;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s.out_diff[o3 + c * plane] = (float)(A[c] * Consts<T>::INV_PI * Dc[c]);
            s.out_spec[o3 + c * plane] = (float)Sc[c];
        }
    }
    // GE[c][u] = dL/d env[c][d0 + u]
    IRB_HD void quad_bwd(const EnvTab& tab, int d0, const T (&E)[3][4], T (&GE)[3][4]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T wx = (T)tab.lx[d0 + u], wy = (T)tab.ly[d0 + u], wz = (T)tab.lz[d0 + u];
            const T om = (T)tab.w[d0 + u];
            if (px.regular) {
                DirL<T> dl;
                dir_fwd_local(dl, px, pl, wx, wy, wz);
                const T w = om * wz;       // n.l = w_d.z: constant, no gradient through it
                T eS = T(0);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    eS += gS[c] * E[c][u];
                    gAlb[c] += w * E[c][u];
                    GE[c][u] = w * (gDa[c] + gS[c] * dl.spec);
                }
                dir_bwd_local(gl, px, pl, dl, wx, wy, wz, w * eS);
                continue;
            }
            Dir<T> dr;
            dir_fwd(dr, px, wx, wy, wz);
            const T w = om * dr.ndl;
            T eD = T(0), eS = T(0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                eD += gDa[c] * E[c][u];
                eS += gS[c] * E[c][u];
                gAlb[c] += w * E[c][u];
                GE[c][u] = w * (gDa[c] + gS[c] * dr.spec);
            }
            dir_bwd(pg, px, dr, wx, wy, wz, om * eD, om * eS);
        }
    }
    IRB_HD void end_bwd(const RenderEnvArgs& s, size_t b, size_t pix, size_t plane) {
        const size_t o3 = (b * 3) * plane + pix;
        if (px.regular) gradl_to_world(pg, gl, px);
        V3<T> g_nt;
        T g_rough;
        pixel_bwd_finish(pg, px, nt, g_nt, g_rough);
        if (s.g_normal) {
            s.g_normal[o3] = (float)g_nt.x;
            s.g_normal[o3 + plane] = (float)g_nt.y;
            s.g_normal[o3 + 2 * plane] = (float)g_nt.z;
        }
        if (s.g_rough) s.g_rough[b * plane + pix] = (float)g_rough;
        if (s.g_albedo) {
#pragma unroll
            for (int c = 0; c < 3; ++c) s.g_albedo[o3 + c * plane] = (float)(gD[c] * Consts<T>::INV_PI * gAlb[c]);
        }
    }
};

template <typename T> IRB_HD void drv_render_env_fwd(const RenderEnvArgs& s, const EnvTab& tab, size_t b, size_t pix) {
    const size_t plane = (size_t)s.R * s.C;
    const int D = tab.D;
    RenderEnvPixel<T> st;
    st.begin(s, b, pix, plane, false);
    for (int d0 = 0; d0 < D; d0 += 4) {
        const F4 e0 = ld4(s.env + ((b * 3 + 0) * plane + pix) * D + d0);
        const F4 e1 = ld4(s.env + ((b * 3 + 1) * plane + pix) * D + d0);
        const F4 e2 = ld4(s.env + ((b * 3 + 2) * plane + pix) * D + d0);
        const T E[3][4] = {{(T)e0.x, (T)e0.y, (T)e0.z, (T)e0.w},
                           {(T)e1.x, (T)e1.y, (T)e1.z, (T)e1.w},
                           {(T)e2.x, (T)e2.y, (T)e2.z, (T)e2.w}};
        st.quad_fwd(tab, d0, E);
    }
    st.end_fwd(s, b, pix, plane);
}

template <typename T> IRB_HD void drv_render_env_bwd(const RenderEnvArgs& s, const EnvTab& tab, size_t b, size_t pix) {
    const size_t plane = (size_t)s.R * s.C;
    const int D = tab.D;
    RenderEnvPixel<T> st;
    st.begin(s, b, pix, plane, true);
    for (int d0 = 0; d0 < D; d0 += 4) {
        const F4 e0 = ld4(s.env + ((b * 3 + 0) * plane + pix) * D + d0);
        const F4 e1 = ld4(s.env + ((b * 3 + 1) * plane + pix) * D + d0);
        const F4 e2 = ld4(s.env + ((b * 3 + 2) * plane + pix) * D + d0);
        const T E[3][4] = {{(T)e0.x, (T)e0.y, (T)e0.z, (T)e0.w},
                           {(T)e1.x, (T)e1.y, (T)e1.z, (T)e1.w},
                           {(T)e2.x, (T)e2.y, (T)e2.z, (T)e2.w}};
        T GE[3][4];
        st.quad_bwd(tab, d0, E, GE);
        if (s.g_env) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                st4(s.g_env + ((b * 3 + c) * plane + pix) * D + d0, (float)GE[c][0], (float)GE[c][1], (float)GE[c][2],
                    (float)GE[c][3]);
        }
    }
    st.end_bwd(s, b, pix, plane);
}

}  // namespace irb
