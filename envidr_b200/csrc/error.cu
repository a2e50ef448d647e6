// error.cu -- thread-local last-error string + version for libenvidr_b200.
#include <stdarg.h>
#include <mutex>
#include "common.cuh"

namespace envidr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Scratch for entry points whose reference signature carries no workspace argument (march_rays_train's chained scan, the replay
// march's block totals): per device ONE allocation of kScratchSlots fixed-size slots per kind, made on first use and never freed,
// moved or grown; every stream gets its own slot (assigned in order of first appearance, no allocation -- so a CUDA-graph capture on a
// fresh side stream works as long as one warm-up call ran on the device before).  A captured graph may therefore bake the pointer
// in, two streams / two devices of one process never share status words, and a request beyond the fixed size is an error instead
// of a reallocation.
constexpr int kScratchSlots = 8, kScratchKinds = 2, kScratchMaxDev = 16;
struct DevScratch { uint8_t* base[kScratchKinds] = {nullptr, nullptr}; cudaStream_t owner[kScratchSlots] = {}; int used = 0; };
static std::mutex g_scratch_mu;
static DevScratch g_scratch[kScratchMaxDev];
void* stream_scratch(int kind, size_t bytes, size_t fixed_bytes, cudaStream_t st) {
    if (kind < 0 || kind >= kScratchKinds || bytes > fixed_bytes) {
        set_error("scratch request %zu B exceeds the fixed %zu B buffer of kind %d", bytes, fixed_bytes, kind);
        return nullptr;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kScratchMaxDev) { set_error("scratch: device index %d out of range", dev); return nullptr; }
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    DevScratch& d = g_scratch[dev];
    if (!d.base[kind]) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, fixed_bytes * kScratchSlots);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("scratch allocation of %zu B failed (%s); inside a CUDA-graph capture run one warm-up call on this device first",
                      fixed_bytes * kScratchSlots, cudaGetErrorString(e));
            return nullptr;
        }
        d.base[kind] = static_cast<uint8_t*>(p);
    }
    int slot = -1;
    for (int i = 0; i < d.used; i++) if (d.owner[i] == st) { slot = i; break; }
    if (slot < 0) {
        if (d.used == kScratchSlots) { set_error("scratch: more than %d streams use this entry point on device %d", kScratchSlots, dev); return nullptr; }
        slot = d.used++;
        d.owner[slot] = st;
    }
    return d.base[kind] + (size_t)slot * fixed_bytes;
}
}  // namespace envidr

extern "C" {
const char* envidr_last_error(void) { return envidr::g_err; }
int envidr_version(void) { return 103; }   // 1.03: + envidr_env_mlp_* / envidr_mlp_* (fused training kernels); 1.02: envidr_get_scatter_idx takes M, NeuS field, sweep shading
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
