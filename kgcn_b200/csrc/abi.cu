// ABI bookkeeping: version and the thread-local last-error string (include/kgcn_b200.h).
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace kgcn {

char* error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_PDL");
        return e == nullptr || e[0] != '0';
    }();
    return on;
}

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace kgcn

extern "C" uint64_t kgcn_launch_count(void) { return kgcn::g_launches.load(std::memory_order_relaxed); }
extern "C" int kgcn_abi_version(void) { return KGCN_B200_ABI_VERSION; }
extern "C" const char* kgcn_last_error(void) { return kgcn::error_buffer(); }
