// Nash solver of the double-oracle level sampler (reference environments/nash_sampler.py:24-58 get_nash,
// util/projection.py:9-38 projection_simplex): projected gradient descent-ascent on the bilinear game
// x^T G y with both strategies projected onto the simplex over their first nz coordinates, averaged
// iterates.  One CTA (the problem is a strictly sequential 10,000-step recursion over <= 1024 levels):
// strategies, sort buffers and running sums live in shared memory, G (<= 4 MB) stays in L2.
#include "common.cuh"
#include "../../include/toued.h"

constexpr int NS_T = 1024;

__device__ void project_simplex(float* v, int n, int nz, float* sk, int* si, float* cs, float* red, int* cnt) {
    // v[0..n) -> projection onto { w >= 0, sum w = 1, w[i >= nz] = 0 } (util/projection.py semantics)
    const int tid = threadIdx.x;
    int P2 = 1; while (P2 < n) P2 <<= 1;
    for (int i = tid; i < P2; i += NS_T) { sk[i] = (i < nz) ? v[i] : -INFINITY; si[i] = i; }
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1)                        // bitonic sort, descending by value (ties: index)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += NS_T) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = sk[i], b = sk[ixj];
                    const int ia = si[i], ib = si[ixj];
                    const bool a_first = (a > b) || (a == b && ia < ib);
                    const bool up = (i & k) == 0;
                    if (a_first != up) { sk[i] = b; sk[ixj] = a; si[i] = ib; si[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    // inclusive cumsum of the sorted values (Hillis-Steele; n <= 1024 -> one element per thread)
    float val = (tid < nz) ? sk[tid] : 0.0f;
    cs[tid] = val;
    __syncthreads();
    for (int d = 1; d < NS_T; d <<= 1) {
        const float add = (tid >= d) ? cs[tid - d] : 0.0f;
        __syncthreads();
        cs[tid] += add;
        __syncthreads();
    }
    if (tid == 0) *cnt = 0;
    __syncthreads();
    if (tid < nz) {
        const float ind = (float)(tid + 1);
        const float c = 1.0f / ind + (sk[tid] - cs[tid] / ind);
        if (c > 0.0f) atomicAdd(cnt, 1);
    }
    __syncthreads();
    const int kk = *cnt;
    const float theta = 1.0f / (float)kk - cs[kk - 1] / (float)kk;
    __syncthreads();
    for (int i = tid; i < n; i += NS_T) v[i] = 0.0f;
    __syncthreads();
    if (tid < nz) v[si[tid]] = fmaxf(sk[tid] + theta, 0.0f);
    __syncthreads();
}

__global__ void __launch_bounds__(NS_T, 1)
get_nash_kernel(const float* __restrict__ G, const float* __restrict__ x0, const float* __restrict__ y0,
                float* __restrict__ x_out, float* __restrict__ y_out, int n, int x_nz, int y_nz, int iters, float lr) {
    __shared__ float x[NS_T], y[NS_T], xn[NS_T], yn[NS_T], xs[NS_T], ys[NS_T], sk[NS_T], cs[NS_T], red[32];
    __shared__ int si[NS_T], cnt;
    const int tid = threadIdx.x;
    x[tid] = tid < n ? x0[tid] : 0.0f; y[tid] = tid < n ? y0[tid] : 0.0f;
    xs[tid] = x[tid]; ys[tid] = y[tid];
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        // x_grad = G y ; y_grad = -(x^T G)   (both with the OLD x, y)
        float gx = 0.0f, gy = 0.0f;
        if (tid < n) {
            for (int j = 0; j < n; ++j) gx = fmaf(G[(size_t)tid * n + j], y[j], gx);
            for (int i = 0; i < n; ++i) gy = fmaf(x[i], G[(size_t)i * n + tid], gy);
        }
        xn[tid] = tid < n ? x[tid] - lr * gx : 0.0f;
        yn[tid] = tid < n ? y[tid] + lr * gy : 0.0f;
        __syncthreads();
        project_simplex(xn, n, x_nz, sk, si, cs, red, &cnt);
        project_simplex(yn, n, y_nz, sk, si, cs, red, &cnt);
        x[tid] = xn[tid]; y[tid] = yn[tid];
        xs[tid] += xn[tid]; ys[tid] += yn[tid];
        __syncthreads();
    }
    if (tid < n) { x_out[tid] = xs[tid] / (float)(iters + 1); y_out[tid] = ys[tid] / (float)(iters + 1); }
}

__global__ void __launch_bounds__(NS_T, 1)
project_simplex_kernel(float* v, int n, int nz) {
    __shared__ float w[NS_T], sk[NS_T], cs[NS_T], red[32];
    __shared__ int si[NS_T], cnt;
    w[threadIdx.x] = threadIdx.x < n ? v[threadIdx.x] : 0.0f;
    __syncthreads();
    project_simplex(w, n, nz, sk, si, cs, red, &cnt);
    if (threadIdx.x < n) v[threadIdx.x] = w[threadIdx.x];
}

extern "C" int toued_get_nash(const float* game, const float* x0, const float* y0, float* x_out, float* y_out,
                              int n, int x_nz, int y_nz, int num_iters, float lr, void* stream) {
    TOUED_CHECK(n >= 1 && n <= NS_T, "toued_get_nash: buffer_size=%d must be in 1..1024", n);
    TOUED_CHECK(x_nz >= 1 && x_nz <= n && y_nz >= 1 && y_nz <= n, "toued_get_nash: bad support sizes");
    get_nash_kernel<<<1, NS_T, 0, (cudaStream_t)stream>>>(game, x0, y0, x_out, y_out, n, x_nz, y_nz, num_iters, lr);
    TOUED_LAUNCH_CHECK();
    return 0;
}

extern "C" int toued_projection_simplex(float* v, int n, int max_nz, void* stream) {
    TOUED_CHECK(n >= 1 && n <= NS_T && max_nz >= 1 && max_nz <= n, "toued_projection_simplex: bad sizes");
    project_simplex_kernel<<<1, NS_T, 0, (cudaStream_t)stream>>>(v, n, max_nz);
    TOUED_LAUNCH_CHECK();
    return 0;
}
