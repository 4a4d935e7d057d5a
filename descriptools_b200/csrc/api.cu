// api.cu -- library-wide plumbing of the C ABI (include/dtb200.h): version, error strings,
// last-CUDA-error text, kernel-launch accounting.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace dtb {
namespace {
std::atomic<int64_t> g_launches{0};
thread_local char g_last_err[512] = "";
}  // namespace

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- per-kernel timing -------------------------------------------------------------------------
namespace {
struct ProfRec { const char *name; cudaEvent_t e0, e1; };
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
thread_local ProfRec g_prof_cur = {nullptr, nullptr, nullptr};

cudaEvent_t prof_event()
{
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void prof_begin(const char *name, cudaStream_t st)
{
    g_prof_cur.name = nullptr;
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_cur = {name, prof_event(), prof_event()};
    cudaEventRecord(g_prof_cur.e0, st);
}

void prof_end(cudaStream_t st)
{
    if (!g_prof_cur.name) return;
    cudaEventRecord(g_prof_cur.e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(g_prof_cur);
    g_prof_cur.name = nullptr;
}

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

extern "C" void dtb_profile_enable(int on)
{
    using namespace dtb;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof_recs) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
    g_prof_recs.clear();
    g_prof_on.store(on != 0);
}

extern "C" int64_t dtb_profile_collect(char *buf, int64_t cap)
{
    using namespace dtb;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<double, int64_t>> agg;
    std::vector<std::string> order;
    for (auto &r : g_prof_recs) {
        float ms = 0.0f;
        cudaEventSynchronize(r.e1);
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); ms = 0.0f; }
        auto it = agg.find(r.name);
        if (it == agg.end()) { agg[r.name] = {ms, 1}; order.push_back(r.name); }
        else { it->second.first += ms; it->second.second += 1; }
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    std::string out;
    char line[256];
    for (auto &n : order) {
        snprintf(line, sizeof(line), "%s %.6f %lld\n", n.c_str(), agg[n].first, (long long)agg[n].second);
        out += line;
    }
    if (buf && cap > 0) {
        const size_t k = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
        memcpy(buf, out.data(), k);
        buf[k] = 0;
    }
    return (int64_t)out.size();
}
