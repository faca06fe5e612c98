// Key derivation on the device: the per-agent key chains of a meta-step.
//
// The reference derives every agent's rollout keys with chains of jax.random.split on the accelerator
// (meta/train.py:38-42,109; agents/lpg_agent.py:104-105; agents/agents.py:99-103).  Doing the same on the
// host costs a threefry pass + an H2D copy in front of every rollout launch; these two kernels keep the
// whole derivation on the GPU, bit-identical to jax 0.4.13's threefry2x32 split (common.cuh).
#include "common.cuh"
#include "../../include/toued.h"

// out[i][j] = jax.random.split(keys_in[i], num)[offset + j],  j < count
__global__ void key_split_kernel(const uint32_t* __restrict__ keys_in, int n_keys, uint32_t num, uint32_t offset,
                                 uint32_t count, uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_keys * count) return;
    const size_t k = i / count;
    const uint32_t j = (uint32_t)(i % count);
    Key key; key.a = keys_in[2 * k]; key.b = keys_in[2 * k + 1];
    const Key r = split_n(key, num, offset + j);
    out[2 * i] = r.a; out[2 * i + 1] = r.b;
}

// for k < K: (carry, keys_out[k][i]) = jax.random.split(carry, 2), carry_0 = keys_in[i]
__global__ void key_chain_kernel(const uint32_t* __restrict__ keys_in, int n, int K, uint32_t* __restrict__ keys_out,
                                 uint32_t* __restrict__ carry_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Key c; c.a = keys_in[2 * i]; c.b = keys_in[2 * i + 1];
    for (int k = 0; k < K; ++k) {
        Key first, second;
        split2(c, first, second);
        keys_out[((size_t)k * n + i) * 2] = second.a;
        keys_out[((size_t)k * n + i) * 2 + 1] = second.b;
        c = first;
    }
    if (carry_out) { carry_out[2 * i] = c.a; carry_out[2 * i + 1] = c.b; }
}

extern "C" int toued_key_split(const uint32_t* keys_in, int n_keys, int num, int offset, int count, uint32_t* out,
                               void* stream) {
    TOUED_CHECK(n_keys > 0 && num > 0 && count > 0 && offset >= 0 && offset + count <= num,
                "toued_key_split: bad range (num=%d offset=%d count=%d)", num, offset, count);
    const size_t total = (size_t)n_keys * count;
    key_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        keys_in, n_keys, (uint32_t)num, (uint32_t)offset, (uint32_t)count, out);
    TOUED_LAUNCH_CHECK();
    return 0;
}

extern "C" int toued_key_chain(const uint32_t* keys_in, int n_keys, int chain_len, uint32_t* keys_out,
                               uint32_t* carry_out, void* stream) {
    TOUED_CHECK(n_keys > 0 && chain_len > 0, "toued_key_chain: empty problem");
    key_chain_kernel<<<(n_keys + 127) / 128, 128, 0, (cudaStream_t)stream>>>(keys_in, n_keys, chain_len, keys_out, carry_out);
    TOUED_LAUNCH_CHECK();
    return 0;
}
