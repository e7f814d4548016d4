// irb_api.cu -- error state, ABI version and the hemisphere direction table.
#include <math.h>
#include <atomic>
#include <string.h>

#include "irb_common.cuh"

namespace irb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Az_j = ((j+.5)/W - .5) 2pi, El_i = ((i+.5)/H) pi/2, ls = (sinEl cosAz, sinEl sinAz, cosEl),
// envWeight = sinEl pi^2/(W H); fp64 then rounded to fp32 like the reference's numpy code
// (models.py:353-362 and 437-448).
bool make_env_tab(EnvTab& tab, int envH, int envW) {
    if (envH <= 0 || envW <= 0 || envH * envW > IRB_MAX_DIRS || (envH * envW) % 4 != 0) return false;
    const double PI = 3.14159265358979323846;
    memset(&tab, 0, sizeof(tab));
    for (int i = 0; i < envH; ++i) {
        const double El = ((i + 0.5) / envH) * PI / 2.0;
        for (int j = 0; j < envW; ++j) {
            const double Az = ((j + 0.5) / envW - 0.5) * 2 * PI;
            const int d = i * envW + j;
            tab.lx[d] = (float)(sin(El) * cos(Az));
            tab.ly[d] = (float)(sin(El) * sin(Az));
            tab.lz[d] = (float)cos(El);
            tab.w[d] = (float)(sin(El) * PI * PI / envW / envH);
        }
    }
    tab.D = envH * envW;
    return true;
}

}  // namespace irb

extern "C" {
int irb_abi_version(void) { return IRB_ABI_VERSION; }
const char* irb_last_error(void) { return irb::g_err; }
long long irb_launch_count(void) { return irb::g_launches.load(std::memory_order_relaxed); }
}
