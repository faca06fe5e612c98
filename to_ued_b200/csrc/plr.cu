// Prioritised Level Replay: the index work of the level buffer on the device (reference
// environments/level_sampler.py:183-234 buffer update / replay-vs-random selection, :331-353 _reset_lowest_scoring,
// :355-387 _replay_from_buffer with score_transform = "rank", :389-408 _sample_random_from_buffer).
//
// The buffer (score f32[B], active u8[B], new u8[B], B <= 8192) and the per-agent inputs (old buffer ids, terminated
// flags, regret scores of the GLOBAL batch, n <= 8192) live in device memory; the regret scores come straight from the
// evaluation kernels.  One CTA does everything: the work is a handful of sorts of <= 8192 keys.  Every choice is an
// integer decision and must be bit-exact with the oracle (oracle/level_sampler.py):
//   * stable argsort = bitonic sort of unique 64-bit keys (order-preserving image of the float << 32 | index);
//   * .at[old_ids].set(...) with duplicate ids = last writer wins (XLA applies scatter updates in order): the last agent
//     index per id is found with an integer atomicMax, then exactly that agent writes;
//   * jax.random draws: the same threefry derivations as everywhere else (common.cuh): split, random_bits, uniform,
//     the sort-based permutation (_shuffle) and the Gumbel top-k over a 0/1 mask (ranked on the raw mantissa bits);
//   * float contract of the rank transform: exp_portable on score / temperature clamped to [-80, 80], a LEFT-TO-RIGHT
//     fp32 sum, IEEE division -- the only floats that can create or break ties.
#include "common.cuh"
#include "../../include/toued.h"

constexpr int PLR_THREADS = 1024;
constexpr int PLR_MAX = 8192;

__device__ __forceinline__ uint32_t plr_orderable(float f) {
    if (f != f) return 0xFFFFFFFFu;                              // NaN sorts last (numpy / XLA sort order)
    if (f == 0.0f) return 0x80000000u;                           // -0.0 and +0.0 compare equal (ties go by index)
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ascending bitonic sort of P (power of two) unique keys in shared memory
__device__ void plr_sort(uint64_t* keys, int P) {
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += PLR_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = keys[i], b = keys[ixj];
                    if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ int plr_pow2(int n) { int p = 2; while (p < n) p <<= 1; return p; }

// ------------------------------------------------------------------------------------------------
// _reset_lowest_scoring (level_sampler.py:331-353): ids of the minimum_new lowest-scoring levels (new levels first:
// -inf; active levels last: +inf), then score <- 0, active <- False at those ids and, quirk Q3, new <- active.at[ids].set(True)
__global__ void __launch_bounds__(PLR_THREADS)
plr_reset_kernel(float* __restrict__ score, uint8_t* __restrict__ active, uint8_t* __restrict__ is_new, int B, int minimum_new,
                 int* __restrict__ reset_ids) {
    extern __shared__ uint64_t keys[];
    const int P = plr_pow2(B);
    for (int i = threadIdx.x; i < P; i += PLR_THREADS) {
        uint64_t k = ~0ull;
        if (i < B) {
            const float s = active[i] ? __int_as_float(0x7F800000) : (is_new[i] ? __int_as_float(0xFF800000) : score[i]);
            k = ((uint64_t)plr_orderable(s) << 32) | (uint32_t)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    plr_sort(keys, P);
    for (int i = threadIdx.x; i < B; i += PLR_THREADS) is_new[i] = active[i];           // Q3: new is rebuilt from active
    __syncthreads();
    for (int j = threadIdx.x; j < minimum_new; j += PLR_THREADS) {
        const int id = (int)(uint32_t)keys[j];
        reset_ids[j] = id;
        score[id] = 0.0f; active[id] = 0; is_new[id] = 1;
    }
}

extern "C" int toued_plr_reset_lowest(float* score, uint8_t* active, uint8_t* is_new, int buffer_size, int minimum_new,
                                      int* reset_ids, void* stream) {
    TOUED_CHECK(buffer_size > 0 && buffer_size <= PLR_MAX, "toued_plr_reset_lowest: buffer_size=%d must be in [1, %d]", buffer_size, PLR_MAX);
    TOUED_CHECK(minimum_new >= 0 && minimum_new <= buffer_size, "toued_plr_reset_lowest: minimum_new=%d exceeds the buffer", minimum_new);
    int P = 2; while (P < buffer_size) P <<= 1;
    const size_t smem = (size_t)P * sizeof(uint64_t);
    TOUED_CUDA(cudaFuncSetAttribute(plr_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    plr_reset_kernel<<<1, PLR_THREADS, smem, (cudaStream_t)stream>>>(score, active, is_new, buffer_size, minimum_new, reset_ids);
    TOUED_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// level_sampler.py:183-234 from "update level buffer with the scores of terminated agents" to "mark the sampled levels
// active".  key = the sampler's rng right before `rng, replay_rng, random_rng = jax.random.split(rng, 3)`.
__global__ void __launch_bounds__(PLR_THREADS)
plr_select_kernel(Key key, float* __restrict__ score, uint8_t* __restrict__ active, uint8_t* __restrict__ is_new, int B,
                  const int* __restrict__ old_ids, const uint8_t* __restrict__ terminated, const float* __restrict__ new_scores,
                  int n, float p_replay, float temperature, int shuffle_rounds, int* __restrict__ new_ids) {
    extern __shared__ uint64_t keys[];                       // P keys
    const int P = plr_pow2(B > n ? B : n);
    int* aux = reinterpret_cast<int*>(keys + P);              // [P]: last writer per id, later the permutation
    int* rep = aux + P;                                       // [n] replay ids
    int* rnd = rep + n;                                       // [n] random ids
    float* sval = reinterpret_cast<float*>(rnd + n);          // [B] transformed scores
    __shared__ int invalid_cnt, n_to_replay;
    __shared__ float total;
    const int tid = threadIdx.x;

    // ---- .at[old_ids].set(where(terminated, new, old)): last writer per id; a non-terminated writer restores the old value
    for (int i = tid; i < B; i += PLR_THREADS) aux[i] = -1;
    if (tid == 0) { invalid_cnt = 0; n_to_replay = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += PLR_THREADS) atomicMax(&aux[old_ids[i]], i);
    __syncthreads();
    for (int i = tid; i < n; i += PLR_THREADS) {
        const int id = old_ids[i];
        if (aux[id] == i && terminated[i]) { score[id] = new_scores[i]; active[id] = 0; is_new[id] = 0; }
    }
    __syncthreads();

    // ---- _replay_from_buffer, rank transform: flip(argsort(p))[:n]
    for (int i = tid; i < B; i += PLR_THREADS) {
        const bool invalid = is_new[i] || active[i];
        if (invalid) atomicAdd(&invalid_cnt, 1);
        const float x = fminf(fmaxf(__fdiv_rn(score[i], temperature), -80.0f), 80.0f);
        sval[i] = invalid ? 0.0f : exp_portable(x);
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (int i = 0; i < B; ++i) s = __fadd_rn(s, sval[i]);                           // left-to-right (the contract)
        total = s;
    }
    __syncthreads();
    const bool too_few = B - invalid_cnt < n;                                            // p_replay <- ones
    for (int i = tid; i < P; i += PLR_THREADS) {
        uint64_t k = 0ull;                                                               // padding sorts first
        if (i < B) k = ((uint64_t)plr_orderable(too_few ? 1.0f : __fdiv_rn(sval[i], total)) << 32) | (uint32_t)i;
        keys[i] = k;
    }
    __syncthreads();
    plr_sort(keys, P);
    for (int j = tid; j < n; j += PLR_THREADS) rep[j] = (int)(uint32_t)keys[P - 1 - j];
    __syncthreads();

    // ---- key plumbing of level_sampler.py:201-226
    Key rng1 = split_n(key, 3, 0), random_rng = split_n(key, 3, 2);                       // (replay_rng is unused by "rank")
    Key rng2, k_unif, rng3, k_perm;
    split2(rng1, rng2, k_unif);
    split2(rng2, rng3, k_perm);

    // ---- _sample_random_from_buffer: choice(..., replace=False, p = new & ~active) = decreasing mantissa, ties by index
    for (int i = tid; i < P; i += PLR_THREADS) {
        uint64_t k = ~0ull;
        if (i < B) {
            const uint32_t mant = bits_elem(random_rng, (uint32_t)B, (uint32_t)i) >> 9;
            const bool ok = is_new[i] && !active[i];
            k = ((uint64_t)(ok ? 0x7FFFFFu - mant : 0x1000000u) << 32) | (uint32_t)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    plr_sort(keys, P);
    for (int j = tid; j < n; j += PLR_THREADS) rnd[j] = (int)(uint32_t)keys[j];

    // ---- n_to_replay = sum(uniform(rng, (n,)) < p_replay)
    int cnt = 0;
    for (int i = tid; i < n; i += PLR_THREADS) cnt += bits_to_unit(bits_elem(k_unif, (uint32_t)n, (uint32_t)i)) < p_replay ? 1 : 0;
    if (cnt) atomicAdd(&n_to_replay, cnt);
    __syncthreads();

    // ---- jax.random.permutation(rng, use_replay): rounds of stable sorts by fresh random bits (jax _shuffle)
    for (int i = tid; i < n; i += PLR_THREADS) aux[i] = i;
    __syncthreads();
    Key kp = k_perm;
    const int Pn = plr_pow2(n);
    for (int r = 0; r < shuffle_rounds; ++r) {
        Key knext, sub;
        split2(kp, knext, sub);
        kp = knext;
        for (int i = tid; i < Pn; i += PLR_THREADS)
            keys[i] = i < n ? (((uint64_t)bits_elem(sub, (uint32_t)n, (uint32_t)i) << 32) | (uint32_t)i) : ~0ull;
        __syncthreads();
        plr_sort(keys, Pn);
        int v[(PLR_MAX + PLR_THREADS - 1) / PLR_THREADS];
        int c = 0;
        for (int j = tid; j < n; j += PLR_THREADS) v[c++] = aux[(uint32_t)keys[j]];
        __syncthreads();
        c = 0;
        for (int j = tid; j < n; j += PLR_THREADS) aux[j] = v[c++];
        __syncthreads();
    }

    // ---- choose, keep the old id where the agent did not terminate, mark the sampled levels active
    const bool can_replay = B - invalid_cnt >= n;
    for (int j = tid; j < n; j += PLR_THREADS) {
        const bool use = can_replay && aux[j] < n_to_replay;
        const int id = terminated[j] ? (use ? rep[j] : rnd[j]) : old_ids[j];
        new_ids[j] = id;
    }
    __syncthreads();
    for (int j = tid; j < n; j += PLR_THREADS) active[new_ids[j]] = 1;
}

extern "C" int toued_plr_select(uint32_t key0, uint32_t key1, float* score, uint8_t* active, uint8_t* is_new, int buffer_size,
                                const int* old_ids, const uint8_t* terminated, const float* new_scores, int n_agents,
                                float p_replay, float temperature, int shuffle_rounds, int* new_ids, void* stream) {
    TOUED_CHECK(buffer_size > 0 && buffer_size <= PLR_MAX, "toued_plr_select: buffer_size=%d must be in [1, %d]", buffer_size, PLR_MAX);
    TOUED_CHECK(n_agents > 0 && n_agents <= buffer_size, "toued_plr_select: n_agents=%d must be in [1, buffer_size]", n_agents);
    TOUED_CHECK(temperature > 0.0f && shuffle_rounds >= 1, "toued_plr_select: bad temperature / shuffle_rounds");
    int P = 2; while (P < buffer_size) P <<= 1;
    const size_t smem = (size_t)P * (sizeof(uint64_t) + sizeof(int)) + (size_t)n_agents * 2 * sizeof(int) + (size_t)buffer_size * sizeof(float);
    TOUED_CUDA(cudaFuncSetAttribute(plr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Key k; k.a = key0; k.b = key1;
    plr_select_kernel<<<1, PLR_THREADS, smem, (cudaStream_t)stream>>>(k, score, active, is_new, buffer_size, old_ids, terminated,
                                                                     new_scores, n_agents, p_replay, temperature, shuffle_rounds, new_ids);
    TOUED_LAUNCH_CHECK();
    return 0;
}
