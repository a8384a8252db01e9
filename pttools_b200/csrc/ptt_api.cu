// Library-level entry points of the C ABI (include/pttools_b200.h).
#include "ptt_host.cuh"

namespace ptt {
char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace ptt

extern "C" const char* ptt_last_error(void) { return ptt::error_buffer(); }
extern "C" int ptt_version(void) { return 100; }
