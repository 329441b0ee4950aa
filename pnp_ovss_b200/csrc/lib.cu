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
