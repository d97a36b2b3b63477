// Library-level entry points of libmonosdf_b200.so: error text, ABI version, launch counter.
#include "common.cuh"

unsigned long long g_msdf_launches = 0;

static thread_local char g_last_error[512] = "";

void msdf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* msdf_last_error(void) { return g_last_error; }
extern "C" int msdf_abi_version(void) { return MSDF_ABI_VERSION; }
extern "C" unsigned long long msdf_launch_count(void) { return g_msdf_launches; }

// ---- per-launch timing -------------------------------------------------------------------------------------
#include <vector>
namespace {
struct ProfRec { int cls; double work, bytes; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}

int msdf_prof_begin(int cls, double work, cudaStream_t st, double bytes) {
    if (!g_prof_on) return -1;
    ProfRec r{cls, work, bytes, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void msdf_prof_end(int slot, cudaStream_t st) {
    if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}

extern "C" int msdf_profile_enable(int on) {
    g_prof_on = on != 0;
    return MSDF_OK;
}
// Sum of the algorithmic bytes of the recorded launches of class `cls` (call before a resetting msdf_profile_read).
extern "C" int msdf_profile_read_bytes(int cls, double* total_bytes) {
    double b = 0.0;
    for (auto& r : g_prof) if (cls < 256 ? (r.cls & 255) == cls : r.cls == cls) b += r.bytes;
    if (total_bytes) *total_bytes = b;
    return MSDF_OK;
}
// Synchronises, sums the recorded launches of class `cls` (milliseconds, work units, count) and, when reset != 0,
// drops all records.
extern "C" int msdf_profile_read(int cls, double* total_ms, double* total_work, long long* count, int reset) {
    MSDF_CUDA_CALL(cudaDeviceSynchronize());
    double ms = 0.0, work = 0.0; long long n = 0;
    for (auto& r : g_prof) {
        if (cls < 256 ? (r.cls & 255) != cls : r.cls != cls) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; work += r.work; ++n; }
    }
    if (total_ms) *total_ms = ms;
    if (total_work) *total_work = work;
    if (count) *count = n;
    if (reset) {
        for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        g_prof.clear();
    }
    return MSDF_OK;
}
