// error.cu -- thread-local last-error string + version for libenvidr_b200.
#include <stdarg.h>
#include "common.cuh"

namespace envidr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace envidr

extern "C" {
const char* envidr_last_error(void) { return envidr::g_err; }
int envidr_version(void) { return 101; }   // 1.01: + density grid, optimizer, ray generation / loss epilogue entry points
int envidr_abi_sizes(uint32_t out[5]) {
    out[0] = (uint32_t)sizeof(envidr_mlp_layer); out[1] = (uint32_t)sizeof(envidr_field); out[2] = (uint32_t)sizeof(envidr_field_out);
    out[3] = (uint32_t)sizeof(envidr_render_opts); out[4] = (uint32_t)sizeof(envidr_render_out);
    return 0;
}
int envidr_abi_sizes_aux(uint32_t out[5]) {
    out[0] = (uint32_t)sizeof(envidr_density_opts); out[1] = (uint32_t)sizeof(envidr_adam_tensor); out[2] = (uint32_t)sizeof(envidr_loss_in);
    out[3] = (uint32_t)sizeof(envidr_loss_opts); out[4] = (uint32_t)sizeof(envidr_sample_log);
    return 0;
}
}
