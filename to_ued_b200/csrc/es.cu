// OpenES ask / tell on the device for the TA-LPG path (reference meta/train.py:133-227 + evosax==0.1.4
// OpenES / Adam GradientOptimizer [3P-recall]).
//   ask : z ~ N(0, I) for popsize/2 members (jax.random.normal = sqrt(2) erfinv(uniform(-1, 1))),
//         x = mean + sigma * [z; -z], written directly in the pair-adjacent order of meta/train.py:152-158
//         (member 2i = +z_i, member 2i+1 = -z_i).
//   tell: theta_grad = 1/(popsize*sigma) * noise^T fitness with fitness negated (maximize=True), then the
//         evosax Adam step on the mean (beta1 .99, beta2 .999, eps 1e-8, bias correction with gen+1).
#include "common.cuh"
#include "../../include/toued.h"

__global__ void __launch_bounds__(256)
es_ask_kernel(const uint32_t* __restrict__ key, const float* __restrict__ mean, float sigma,
              float* __restrict__ cand, uint32_t half_pop, uint32_t P, uint32_t cs) {
    Key k; k.a = key[0]; k.b = key[1];
    const uint32_t total = half_pop * P;                    // number of normals
    const uint32_t half = (total + 1u) >> 1;
    const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= half) return;
    uint32_t x0 = blk, x1 = half + blk;
    if (x1 >= total) x1 = 0u;
    threefry2x32(k, x0, x1);
    const float lo = -0.99999994f;                          // nextafter(-1, 0)
    const uint32_t idx[2] = {blk, half + blk};
    const uint32_t bits[2] = {x0, x1};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (idx[q] >= total) continue;
        float u = bits_to_unit(bits[q]) * (1.0f - lo) + lo;
        u = fmaxf(lo, u);
        const float z = 1.41421356237309505f * erfinvf(u);
        const uint32_t i = idx[q] / P, p = idx[q] % P;
        const float m = mean[p];
        cand[(size_t)(2 * i) * cs + p] = m + sigma * z;
        cand[(size_t)(2 * i + 1) * cs + p] = m - sigma * z;
    }
}

extern "C" int toued_es_ask(const uint32_t* key, const float* mean, float sigma, float* candidates, int popsize,
                            int n_params, int cand_stride, void* stream) {
    TOUED_CHECK(cand_stride >= n_params, "toued_es_ask: cand_stride < n_params");
    TOUED_CHECK(popsize >= 2 && (popsize & 1) == 0 && n_params > 0, "toued_es_ask: popsize must be even");
    const uint64_t total = (uint64_t)(popsize / 2) * n_params;
    TOUED_CHECK(total < (1ull << 32), "toued_es_ask: too many parameters");
    const uint32_t half = (uint32_t)((total + 1) >> 1);
    es_ask_kernel<<<(half + 255) / 256, 256, 0, (cudaStream_t)stream>>>(key, mean, sigma, candidates, popsize / 2, n_params, cand_stride);
    TOUED_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
es_tell_kernel(const float* __restrict__ cand, const float* __restrict__ fitness, float* __restrict__ mean,
               float* __restrict__ m, float* __restrict__ v, int popsize, int P, int cs, float sigma, float lrate,
               float b1, float b2, float eps, float c1, float c2, float mean_decay) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float mu = mean[p];
    float g = 0.f;
    for (int i = 0; i < popsize; ++i) g = fmaf((cand[(size_t)i * cs + p] - mu) / sigma, -fitness[i], g);
    g *= 1.0f / ((float)popsize * sigma);
    const float mi = (1.0f - b1) * g + b1 * m[p];
    const float vi = (1.0f - b2) * g * g + b2 * v[p];
    m[p] = mi; v[p] = vi;
    float nm = mu - lrate * (mi / c1) / (sqrtf(vi / c2) + eps);
    mean[p] = nm * (1.0f - mean_decay);
}

extern "C" int toued_es_tell(const float* candidates, const float* fitness, float* mean, float* m, float* v,
                             int popsize, int n_params, int cand_stride, float sigma, float lrate, float beta1,
                             float beta2, float eps, int gen_counter, float mean_decay, void* stream) {
    TOUED_CHECK(popsize >= 2 && n_params > 0 && gen_counter >= 0, "toued_es_tell: bad arguments");
    const float c1 = 1.0f - powf(beta1, (float)(gen_counter + 1)), c2 = 1.0f - powf(beta2, (float)(gen_counter + 1));
    es_tell_kernel<<<(n_params + 255) / 256, 256, 0, (cudaStream_t)stream>>>(candidates, fitness, mean, m, v, popsize,
                                                                             n_params, cand_stride, sigma, lrate, beta1, beta2, eps,
                                                                             c1, c2, mean_decay);
    TOUED_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU ES: a rank holds the antithetic pairs [pair_offset, pair_offset + n_pairs) of the
// global population.  ask_shard draws exactly the normals the single-rank ask would draw for those pairs (element
// j = pair * P + p of the global jax.random.normal vector; one threefry block per element, the unused half of the
// block is discarded), grad_partial forms this rank's part of noise^T fitness, and after the all-reduce es_adam
// applies evosax's Adam step -- identical on every rank.
__global__ void __launch_bounds__(256)
es_ask_shard_kernel(const uint32_t* __restrict__ key, const float* __restrict__ mean, float sigma,
                    float* __restrict__ cand, uint32_t half_pop, uint32_t P, uint32_t cs, uint32_t pair_offset,
                    uint32_t n_pairs) {
    Key k; k.a = key[0]; k.b = key[1];
    const uint32_t total = half_pop * P, half = (total + 1u) >> 1;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_pairs * P) return;
    const uint32_t i = e / P, p = e % P;
    const uint32_t j = (pair_offset + i) * P + p;             // index into the global normal vector
    const uint32_t blk = j < half ? j : j - half;
    uint32_t x0 = blk, x1 = half + blk;
    if (x1 >= total) x1 = 0u;
    threefry2x32(k, x0, x1);
    const float lo = -0.99999994f;
    float u = bits_to_unit(j < half ? x0 : x1) * (1.0f - lo) + lo;
    u = fmaxf(lo, u);
    const float z = 1.41421356237309505f * erfinvf(u);
    const float m = mean[p];
    cand[(size_t)(2 * i) * cs + p] = m + sigma * z;
    cand[(size_t)(2 * i + 1) * cs + p] = m - sigma * z;
}

extern "C" int toued_es_ask_shard(const uint32_t* key, const float* mean, float sigma, float* candidates,
                                  int popsize_global, int n_params, int cand_stride, int pair_offset, int n_pairs,
                                  void* stream) {
    TOUED_CHECK(cand_stride >= n_params, "toued_es_ask_shard: cand_stride < n_params");
    TOUED_CHECK(popsize_global >= 2 && (popsize_global & 1) == 0 && n_params > 0, "toued_es_ask_shard: popsize must be even");
    TOUED_CHECK(pair_offset >= 0 && n_pairs > 0 && pair_offset + n_pairs <= popsize_global / 2,
                "toued_es_ask_shard: pairs [%d, %d) outside the population", pair_offset, pair_offset + n_pairs);
    const uint64_t total = (uint64_t)(popsize_global / 2) * n_params;
    TOUED_CHECK(total < (1ull << 32), "toued_es_ask_shard: too many parameters");
    const uint64_t local = (uint64_t)n_pairs * n_params;
    es_ask_shard_kernel<<<(unsigned)((local + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        key, mean, sigma, candidates, popsize_global / 2, n_params, cand_stride, pair_offset, n_pairs);
    TOUED_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
es_grad_partial_kernel(const float* __restrict__ cand, const float* __restrict__ fitness, const float* __restrict__ mean,
                       float* __restrict__ g_out, int n_members, int P, int cs, float sigma) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float mu = mean[p];
    float g = 0.f;
    for (int i = 0; i < n_members; ++i) g = fmaf((cand[(size_t)i * cs + p] - mu) / sigma, -fitness[i], g);
    g_out[p] = g;
}

extern "C" int toued_es_grad_partial(const float* candidates, const float* fitness, const float* mean, float* grad_sum,
                                     int n_members, int n_params, int cand_stride, float sigma, void* stream) {
    TOUED_CHECK(n_members >= 2 && n_params > 0, "toued_es_grad_partial: bad arguments");
    es_grad_partial_kernel<<<(n_params + 255) / 256, 256, 0, (cudaStream_t)stream>>>(candidates, fitness, mean, grad_sum,
                                                                                     n_members, n_params, cand_stride, sigma);
    TOUED_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
es_adam_kernel(const float* __restrict__ g_sum, float* __restrict__ mean, float* __restrict__ m, float* __restrict__ v,
               int popsize, int P, float sigma, float lrate, float b1, float b2, float eps, float c1, float c2,
               float mean_decay) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float g = g_sum[p] * (1.0f / ((float)popsize * sigma));
    const float mi = (1.0f - b1) * g + b1 * m[p];
    const float vi = (1.0f - b2) * g * g + b2 * v[p];
    m[p] = mi; v[p] = vi;
    const float nm = mean[p] - lrate * (mi / c1) / (sqrtf(vi / c2) + eps);
    mean[p] = nm * (1.0f - mean_decay);
}

extern "C" int toued_es_adam(const float* grad_sum, float* mean, float* m, float* v, int popsize_global, int n_params,
                             float sigma, float lrate, float beta1, float beta2, float eps, int gen_counter,
                             float mean_decay, void* stream) {
    TOUED_CHECK(popsize_global >= 2 && n_params > 0 && gen_counter >= 0, "toued_es_adam: bad arguments");
    const float c1 = 1.0f - powf(beta1, (float)(gen_counter + 1)), c2 = 1.0f - powf(beta2, (float)(gen_counter + 1));
    es_adam_kernel<<<(n_params + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grad_sum, mean, m, v, popsize_global, n_params,
                                                                             sigma, lrate, beta1, beta2, eps, c1, c2, mean_decay);
    TOUED_LAUNCH_CHECK();
    return 0;
}
