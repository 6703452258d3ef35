// Error plumbing for the C-ABI: entry points return an int code and leave the
// text in a thread-local buffer read back through cs_last_error().
#include "cs_platform.cuh"
#include <cstdarg>

namespace cs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

}  // namespace cs
