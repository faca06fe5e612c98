// Error plumbing + version for the C ABI (include/toued.h).
#include "common.cuh"
#include "../../include/toued.h"
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";

void toued_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* toued_last_error(void) { return g_err; }

extern "C" int toued_version(void) { return 1; }

// Host-side key plumbing (no GPU involved): out[i][0..n) = threefry_2x32(keys[i], iota(n)) with jax 0.4.13's
// padding / half-split rule -- the same bits_elem the kernels use.  The level generator (configs.py) and the
// level sampler draw a few hundred such vectors per meta-step; in numpy that costs ~20 ms on PLR configurations.
extern "C" int toued_host_iota_bits(const uint32_t* keys, int n_keys, int n, uint32_t* out) {
    if (n_keys < 0 || n < 0) return 1;
    for (int i = 0; i < n_keys; ++i) {
        Key k; k.a = keys[2 * i]; k.b = keys[2 * i + 1];
        uint32_t* o = out + (size_t)i * n;
        const uint32_t half = ((uint32_t)n + 1u) >> 1;
        for (uint32_t j = 0; j < half; ++j) {                 // one threefry block gives elements j and half + j
            uint32_t x0 = j, x1 = half + j;
            if (x1 >= (uint32_t)n) x1 = 0u;
            threefry2x32(k, x0, x1);
            o[j] = x0;
            if (half + j < (uint32_t)n) o[half + j] = x1;
        }
    }
    return 0;
}
