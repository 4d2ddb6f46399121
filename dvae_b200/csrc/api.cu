// Error plumbing and version of libdvae_b200.
#include <stdarg.h>

#include "common.cuh"

namespace dvae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // namespace dvae

extern "C" int dvae_version(void) { return DVAE_ABI_VERSION; }
extern "C" const char* dvae_last_error(void) { return dvae::g_err; }
