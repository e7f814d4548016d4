// hostcheck.cpp -- TEST INFRASTRUCTURE.  Instantiates the per-pixel drivers of
// inverserenderingofindoorscene_b200/csrc/*.cuh on the HOST (float and double) so that the
// shading math and the hand-derived backward can be checked against the oracle in a container
// without a GPU.  Never linked into the product library; loaded only by tests/.
#include <vector>
#include "../../inverserenderingofindoorscene_b200/csrc/irb_driver_env.cuh"

using namespace irb;

template <int KC, typename F> static void for_pixels(int B, int R, int C, F f) {
    for (size_t b = 0; b < (size_t)B; ++b)
        for (size_t p = 0; p < (size_t)R * C; ++p) f(b, p);
}

// Host emulation of the warp-cooperative policy of irb_render_sg_group.cuh (WarpCoop): 32 "lanes" each take the direction
// slice (lane, 32); pass 0 records every value a lane hands to reduce(), the logs are summed element-wise, and pass 1
// re-runs lane 0 with reduce() returning the sums and leader() = true.  Exercises DirSlice / Coop::reduce / Coop::leader
// of the shared drivers exactly as the device uses them.
struct HostCoop {
    static thread_local int lane, pass;
    static thread_local std::vector<double>* log;
    static thread_local const std::vector<double>* totals;
    static thread_local size_t cursor;
    static DirSlice slice() { return DirSlice{lane, 32}; }
    template <typename T> static void reduce(T& v) {
        if (pass == 0) log->push_back((double)v);
        else v = (T)(*totals)[cursor++];
    }
    static bool leader() { return pass == 1; }
};
thread_local int HostCoop::lane = 0;
thread_local int HostCoop::pass = 0;
thread_local std::vector<double>* HostCoop::log = nullptr;
thread_local const std::vector<double>* HostCoop::totals = nullptr;
thread_local size_t HostCoop::cursor = 0;

template <typename F> static void run_coop(F body) {
    std::vector<double> tot;
    for (int l = 0; l < 32; ++l) {
        std::vector<double> lg;
        HostCoop::lane = l; HostCoop::pass = 0; HostCoop::log = &lg;
        body();
        if (l == 0) tot.assign(lg.size(), 0.0);
        for (size_t i = 0; i < lg.size(); ++i) tot[i] += lg[i];
    }
    HostCoop::lane = 0; HostCoop::pass = 1; HostCoop::totals = &tot; HostCoop::cursor = 0;
    body();
}

extern "C" {

int hc_render_sg_fwd_coop(const RenderSgArgs* a, const EnvTab* tab) {
    const size_t plane = (size_t)a->R * a->C;
    for (size_t b = 0; b < (size_t)a->B; ++b)
        for (size_t p = 0; p < plane; ++p) {
            const PixView v{b, p, plane};
            run_coop([&] {
                if (a->K <= 12) drv_render_sg_fwd<double, 12, HostCoop>(*a, v, *a, v, *tab);
                else drv_render_sg_fwd<double, 24, HostCoop>(*a, v, *a, v, *tab);
            });
        }
    return 0;
}

int hc_render_sg_bwd_coop(const RenderSgArgs* a, const EnvTab* tab) {
    if (a->K > 24) return -2;
    const size_t plane = (size_t)a->R * a->C;
    for (size_t b = 0; b < (size_t)a->B; ++b)
        for (size_t p = 0; p < plane; ++p) {
            const PixView v{b, p, plane};
            run_coop([&] {
                if (a->K <= 12) drv_render_sg_bwd<double, 12, 12, true, true, HostCoop>(*a, v, *a, v, *tab);
                else drv_render_sg_bwd<double, 24, 12, true, true, HostCoop>(*a, v, *a, v, *tab);
            });
        }
    return 0;
}

int hc_render_sg_fwd(int dbl, const RenderSgArgs* a, const EnvTab* tab) {
    const int kc = a->K <= 12 ? 12 : 24;
    for_pixels<0>(a->B, a->R, a->C, [&](size_t b, size_t p) {
        if (dbl) { if (kc == 12) drv_render_sg_fwd<double, 12>(*a, *tab, b, p); else drv_render_sg_fwd<double, 24>(*a, *tab, b, p); }
        else     { if (kc == 12) drv_render_sg_fwd<float, 12>(*a, *tab, b, p);  else drv_render_sg_fwd<float, 24>(*a, *tab, b, p); }
    });
    return 0;
}

int hc_render_sg_bwd(int dbl, const RenderSgArgs* a, const EnvTab* tab) {
    if (a->K > 24) return -2;
    for_pixels<0>(a->B, a->R, a->C, [&](size_t b, size_t p) {
        if (a->K <= 12) {
            if (dbl) drv_render_sg_bwd<double, 12, 12, true, true>(*a, *tab, b, p);
            else drv_render_sg_bwd<float, 12, 12, true, true>(*a, *tab, b, p);
        } else {
            if (dbl) drv_render_sg_bwd<double, 24, 12, true, true>(*a, *tab, b, p);
            else drv_render_sg_bwd<float, 24, 12, true, true>(*a, *tab, b, p);
        }
    });
    return 0;
}

int hc_sg_to_env_fwd(int dbl, const SgEnvArgs* s, const EnvTab* tab) {
    for_pixels<0>(s->B, s->R, s->C, [&](size_t b, size_t p) {
        if (dbl) drv_sg_to_env_fwd<double, 12>(*s, *tab, b, p); else drv_sg_to_env_fwd<float, 12>(*s, *tab, b, p);
    });
    return 0;
}
int hc_sg_to_env_bwd(int dbl, const SgEnvArgs* s, const EnvTab* tab) {
    for_pixels<0>(s->B, s->R, s->C, [&](size_t b, size_t p) {
        if (dbl) drv_sg_to_env_bwd<double, 12>(*s, *tab, b, p); else drv_sg_to_env_bwd<float, 12>(*s, *tab, b, p);
    });
    return 0;
}
int hc_render_env_fwd(int dbl, const RenderEnvArgs* s, const EnvTab* tab) {
    for_pixels<0>(s->B, s->R, s->C, [&](size_t b, size_t p) {
        if (dbl) drv_render_env_fwd<double>(*s, *tab, b, p); else drv_render_env_fwd<float>(*s, *tab, b, p);
    });
    return 0;
}
int hc_render_env_bwd(int dbl, const RenderEnvArgs* s, const EnvTab* tab) {
    for_pixels<0>(s->B, s->R, s->C, [&](size_t b, size_t p) {
        if (dbl) drv_render_env_bwd<double>(*s, *tab, b, p); else drv_render_env_bwd<float>(*s, *tab, b, p);
    });
    return 0;
}
size_t hc_sizeof_render_sg_args() { return sizeof(RenderSgArgs); }
size_t hc_sizeof_sg_env_args() { return sizeof(SgEnvArgs); }
size_t hc_sizeof_render_env_args() { return sizeof(RenderEnvArgs); }
size_t hc_sizeof_envtab() { return sizeof(EnvTab); }
}
