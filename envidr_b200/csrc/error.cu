// error.cu -- thread-local last-error string + version for libenvidr_b200.
#include <stdarg.h>
#include <map>
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
// march's block totals): ONE fixed-size buffer per (device, stream, kind), allocated on first use and never freed or grown.  A
// captured CUDA graph may therefore bake the pointer in, two streams / two devices of one process never share status words, and a
// request beyond the fixed size is an error instead of a reallocation.
struct ScratchKey { int dev; cudaStream_t st; int kind; bool operator<(const ScratchKey& o) const {
    return dev != o.dev ? dev < o.dev : (st != o.st ? st < o.st : kind < o.kind); } };
static std::mutex g_scratch_mu;
static std::map<ScratchKey, void*> g_scratch;
void* stream_scratch(int kind, size_t bytes, size_t fixed_bytes, cudaStream_t st) {
    if (bytes > fixed_bytes) { set_error("scratch request %zu B exceeds the fixed %zu B buffer of kind %d", bytes, fixed_bytes, kind); return nullptr; }
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    ScratchKey key{dev, st, kind};
    auto it = g_scratch.find(key);
    if (it != g_scratch.end()) return it->second;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, fixed_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("scratch allocation of %zu B failed (%s); inside a CUDA-graph capture run one warm-up call on the capture stream first",
                  fixed_bytes, cudaGetErrorString(e));
        return nullptr;
    }
    g_scratch[key] = p;
    return p;
}
}  // namespace envidr

extern "C" {
const char* envidr_last_error(void) { return envidr::g_err; }
int envidr_version(void) { return 102; }   // 1.02: envidr_get_scatter_idx takes M (bounded writes); + NeuS field, sweep shading entry points
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
