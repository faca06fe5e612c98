// Agent (re-)creation on the device: lecun-normal table initialisation from threefry keys and
// masked environment reset.  Replaces, for the tabular nets, agents/agents.py:31-95 create_agent /
// create_value_critic (flax Dense default kernel_init = lecun_normal = truncated normal on [-2, 2]
// with std sqrt(1/fan_in)/0.87962566, drawn as sqrt(2)*erfinv(uniform(erf(-sqrt2), erf(sqrt2)))
// [3P-recall]) and the batch_reset of environments/level_sampler.py:273-291 _create_agent, which the
// reference runs for every agent on every meta-step and then masks (level_sampler.py:239,264).
#include "gridworld.cuh"
#include "../../include/toued.h"

__global__ void __launch_bounds__(256)
init_tables_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ mask, float* __restrict__ tables,
                   int D, int C, float scale) {
    const int n = blockIdx.y;
    if (mask && !mask[n]) return;
    Key k; k.a = keys[2 * n]; k.b = keys[2 * n + 1];
    const uint32_t total = (uint32_t)D * (uint32_t)C;
    const uint32_t half = (total + 1u) >> 1;
    const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= half) return;
    uint32_t x0 = blk, x1 = half + blk;
    if (x1 >= total) x1 = 0u;
    threefry2x32(k, x0, x1);
    const float lo = -0.95449973610364158f, hi = 0.95449973610364158f;      // erf(-+ 2/sqrt(2))
    float* t = tables + (size_t)n * D * 8;
    const uint32_t idx[2] = {blk, half + blk};
    const uint32_t bits[2] = {x0, x1};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (idx[q] >= total) continue;
        float u = bits_to_unit(bits[q]) * (hi - lo) + lo;                     // jax uniform(minval, maxval)
        u = fmaxf(lo, u);
        float v = 1.41421356237309505f * erfinvf(u);
        v = fminf(fmaxf(v, -1.99999988f), 1.99999988f);                       // clip to (nextafter(-2), nextafter(2))
        const uint32_t d = idx[q] / (uint32_t)C, c = idx[q] % (uint32_t)C;
        t[(size_t)d * 8 + c] = v * scale;
    }
}

__global__ void zero_pad_kernel(const uint8_t* __restrict__ mask, float* __restrict__ tables, int D, int C) {
    const int n = blockIdx.y;
    if (mask && !mask[n]) return;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float* t = tables + ((size_t)n * D + d) * 8;
    for (int c = C; c < 8; ++c) t[c] = 0.0f;
}

extern "C" int toued_init_tables(const uint32_t* keys, const uint8_t* mask, float* tables, int n_agents,
                                 int obs_dim, int n_out, void* stream) {
    TOUED_CHECK(n_agents > 0 && obs_dim > 0 && n_out >= 1 && n_out <= 8, "toued_init_tables: bad shape");
    const uint32_t half = ((uint32_t)obs_dim * n_out + 1u) >> 1;
    const float scale = sqrtf(1.0f / (float)obs_dim) / 0.87962566103423978f;
    cudaStream_t st = (cudaStream_t)stream;
    init_tables_kernel<<<dim3((half + 255) / 256, n_agents), 256, 0, st>>>(keys, mask, tables, obs_dim, n_out, scale);
    TOUED_LAUNCH_CHECK();
    if (n_out < 8) {
        zero_pad_kernel<<<dim3((obs_dim + 255) / 256, n_agents), 256, 0, st>>>(mask, tables, obs_dim, n_out);
        TOUED_LAUNCH_CHECK();
    }
    return 0;
}

// masked reset of N x W environments + step counters (agents whose lifetime ended get a new level)
__global__ void masked_reset_kernel(const LevelRec* __restrict__ levels, const uint8_t* __restrict__ mask,
                                    int32_t* __restrict__ state, int32_t* __restrict__ obs,
                                    int32_t* __restrict__ step, int n_envs, int W, int G2) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_envs) return;
    const int n = g / W;
    if (mask && !mask[n]) return;
    const EnvRegs s = env_reset(levels[n]);
    state[g] = pack_state(s.pos, s.exists, s.time);
    if (obs) obs[g] = pack_obs(obs_row(s, G2), s.time);
    if (step && g % W == 0) step[n] = 0;
}

extern "C" int toued_masked_reset(const void* levels, const uint8_t* mask, int32_t* state, int32_t* obs,
                                  int32_t* step, int n_agents, int n_workers, int max_grid_size, void* stream) {
    TOUED_CHECK(n_agents > 0 && n_workers > 0, "toued_masked_reset: empty problem");
    const int n = n_agents * n_workers;
    masked_reset_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        (const LevelRec*)levels, mask, state, obs, step, n, n_workers, max_grid_size * max_grid_size);
    TOUED_LAUNCH_CHECK();
    return 0;
}
