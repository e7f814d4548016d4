// irb_bilateral.cu -- bilateral-grid solver (placeholder until the kernels land this round).
#include "irb_common.cuh"
using namespace irb;
extern "C" size_t irb_bg_workspace_bytes(int, int, int, int) { return 0; }
extern "C" int irb_bg_build(const float*, int, int, int, double, double, double, void*, size_t, int*, void*) {
    set_error("irb_bg_build: not implemented yet");
    return IRB_E_UNSUPPORTED;
}
extern "C" int irb_bg_vertex_index(const void*, int32_t*, int, int, int, void*) {
    set_error("irb_bg_vertex_index: not implemented yet");
    return IRB_E_UNSUPPORTED;
}
extern "C" int irb_bg_solve_fwd(void*, const float*, const float*, int, int, int, int, double, double, double, int, float*,
                                double*, void*) {
    set_error("irb_bg_solve_fwd: not implemented yet");
    return IRB_E_UNSUPPORTED;
}
extern "C" int irb_bg_solve_bwd(void*, const float*, const float*, const float*, const double*, int, int, int, int, double,
                                double, double, int, float*, float*, void*) {
    set_error("irb_bg_solve_bwd: not implemented yet");
    return IRB_E_UNSUPPORTED;
}
