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
