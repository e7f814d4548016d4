// irb_common.cuh -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/irb.h"
#include "irb_pixel.cuh"

namespace irb {

void set_error(const char* fmt, ...);
// hemisphere table (models.py:353-362, 437-448); returns false if unsupported
bool make_env_tab(EnvTab& tab, int envH, int envW);

// every kernel launch of the library is counted (irb_launch_count): n = launches issued since the last check
void count_launches(int n);
inline int check_launch(const char* what, int n = 1) {
    count_launches(n);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // namespace irb

#define IRB_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            irb::set_error(__VA_ARGS__);       \
            return IRB_E_ARG;                  \
        }                                      \
    } while (0)
