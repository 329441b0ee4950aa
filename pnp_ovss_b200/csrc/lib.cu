// lib.cu -- ABI bookkeeping for libpnp_ovss_b200.so.
#include "common.cuh"

extern "C" int pnp_abi_version(void) { return PNP_ABI_VERSION; }

extern "C" int pnp_compiled_sm(void) { return 100; }

extern "C" const char *pnp_error_string(int code) {
    if (code == PNP_OK) return "ok";
    if (code == PNP_ERR_INVALID_ARGUMENT) return "invalid argument (shape, null pointer, alignment or unsupported size)";
    if (code == PNP_ERR_WORKSPACE) return "workspace too small";
    if (code <= PNP_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(PNP_ERR_CUDA_BASE - code));
    return "unknown error";
}

// ------------------------------------------------------------------------------------------ in-situ kernel timing
#include <vector>

namespace pnp {
namespace prof {
unsigned g_mask = 0;
bool g_filter = false;
cudaStream_t g_filter_stream = nullptr;
struct Span { int id; cudaEvent_t a, b; };
static std::vector<Span> g_spans;   // recorded this session
static std::vector<Span> g_pool;    // reusable event pairs

void begin(int id, cudaStream_t st) {
    Span s;
    if (!g_pool.empty()) {
        s = g_pool.back();
        g_pool.pop_back();
    } else {
        if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
    }
    s.id = id;
    cudaEventRecord(s.a, st);
    g_spans.push_back(s);
}
void end(int id, cudaStream_t st) {
    // g_spans may have reallocated since begin(); the open span of `id` is the last one with that id
    for (size_t i = g_spans.size(); i-- > 0;)
        if (g_spans[i].id == id) { cudaEventRecord(g_spans[i].b, st); break; }
}
}  // namespace prof
}  // namespace pnp

static const char *kKernelNames[pnp::kNumKernelIds] = {
    "", "softmax_fwd", "softmax_bwd_gradcam", "token_merge", "salience_dropout_round", "threshold_prep", "upsample_write",
    "blur_vertical", "blur_horizontal", "blur_normalize", "lattice_build", "crf_unary", "crf_splat_bilateral",
    "crf_blur_axis_bilateral", "crf_meanfield_update", "argmax_channels", "confusion", "crf_splat_spatial",
    "crf_blur_axis_spatial", "tf32_split3", "gelu_tf32_split3", "layernorm_tf32_split3", "lowrank_blur", "lowrank_unary",
    "background_blur", "attention_fp16x3"};

extern "C" const char *pnp_profile_kernel_name(int kernel_id) {
    return (kernel_id > 0 && kernel_id < pnp::kNumKernelIds) ? kKernelNames[kernel_id] : "";
}

extern "C" int pnp_profile_num_kernels(void) { return pnp::kNumKernelIds; }

extern "C" int pnp_profile_start(unsigned kernel_mask) {
    using namespace pnp::prof;
    for (auto &s : g_spans) g_pool.push_back(s);
    g_spans.clear();
    g_spans.reserve(1 << 14);
    g_mask = kernel_mask;
    return PNP_OK;
}

extern "C" int pnp_profile_stop(float *total_ms, int *n_launches, int n_ids) {
    using namespace pnp::prof;
    g_mask = 0;
    for (int i = 0; i < n_ids; ++i) {
        if (total_ms) total_ms[i] = 0.f;
        if (n_launches) n_launches[i] = 0;
    }
    for (auto &s : g_spans) {
        cudaError_t e = cudaEventSynchronize(s.b);
        if (e != cudaSuccess) return pnp::cuda_err(e);
        float ms = 0.f;
        e = cudaEventElapsedTime(&ms, s.a, s.b);
        if (e != cudaSuccess) return pnp::cuda_err(e);
        if (s.id < n_ids) {
            if (total_ms) total_ms[s.id] += ms;
            if (n_launches) n_launches[s.id] += 1;
        }
    }
    return PNP_OK;
}

extern "C" int pnp_profile_filter_stream(pnp_stream_t stream, int enabled) {
    pnp::prof::g_filter = enabled != 0;
    pnp::prof::g_filter_stream = pnp::as_stream(stream);
    return PNP_OK;
}
