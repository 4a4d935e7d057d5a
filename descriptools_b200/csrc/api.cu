// api.cu -- library-wide plumbing of the C ABI (include/dtb200.h): version, error strings,
// last-CUDA-error text, kernel-launch accounting.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace dtb {
namespace {
std::atomic<int64_t> g_launches{0};
thread_local char g_last_err[512] = "";
}  // namespace

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    cudaGetLastError();  // clear the sticky-less error state
    return DTB_ERR_CUDA;
}

int cuda_fail_msg(const char *what)
{
    snprintf(g_last_err, sizeof(g_last_err), "%s", what);
    return DTB_ERR_CUDA;
}
}  // namespace dtb

extern "C" int dtb_abi_version(void) { return DTB_ABI_VERSION; }

extern "C" const char *dtb_error_string(int code)
{
    switch (code) {
    case DTB_OK: return "ok";
    case DTB_ERR_INVALID: return "invalid argument";
    case DTB_ERR_CUDA: return "CUDA error (see dtb_last_cuda_error)";
    case DTB_ERR_WORKSPACE: return "workspace too small";
    case DTB_ERR_UNSUPPORTED: return "unsupported size or parameter";
    default: return "unknown error";
    }
}

extern "C" const char *dtb_last_cuda_error(void) { return dtb::g_last_err; }
extern "C" int64_t dtb_launch_count(void) { return dtb::g_launches.load(std::memory_order_relaxed); }
extern "C" void dtb_reset_launch_count(void) { dtb::g_launches.store(0, std::memory_order_relaxed); }
