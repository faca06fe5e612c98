// Shared device helpers for the TO-UED / GROOVE hot path on sm_100a.
//
//  * threefry2x32 + the jax==0.4.13 key-derivation rules (split / random_bits / uniform) that the
//    reference's hot path uses (environments/rollout.py:40,49,61-64; gridworld.py:76,88,116).
//    The CPU oracle (oracle/prng.py) implements the same contract; every draw is bit-comparable.
//  * exp_portable: exp() built only from individually rounded f32 mul/add + integer exponent
//    arithmetic so the policy softmax that feeds action sampling is bit-identical on CPU and GPU.
//  * LevelRec: one gridworld level (reference EnvParams, gridworld.py:22-35) packed into 192 B so a
//    block can stage its levels in shared memory with one cp.async.bulk (TMA 1-D bulk copy).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TOUED_MAX_OBJS 8
#define TOUED_ACT_PAD 8      // actor rows are padded 5 -> 8 floats (one 32 B sector per row gather)
#define TOUED_NUM_ACTIONS 5  // gridworld.py:219-222

// ------------------------------------------------------------------------------------------------
// Level record.  Per-object tables are pre-gathered with obj_ids at upload time
// (jnp.take(params.obj_X, params.obj_ids), gridworld.py:87,115,122; negative ids wrap).
struct __align__(16) LevelRec {
    int32_t max_steps;                    // max_steps_in_episode
    int32_t grid_size;
    int32_t start_pos;
    int32_t n_objs;
    int32_t lifetime;                     // util/data.py:46-50 Level.lifetime
    int32_t buffer_id;                    // Level.buffer_id
    int32_t _pad0[2];
    int32_t obj_pos[TOUED_MAX_OBJS];      // static_obj_poss
    float obj_reward[TOUED_MAX_OBJS];
    float obj_p_term[TOUED_MAX_OBJS];
    float obj_p_resp[TOUED_MAX_OBJS];
    uint32_t walls[8];                    // bitmask over max_grid_size^2 (<= 256) cells
};
static_assert(sizeof(LevelRec) == 192, "LevelRec must stay 192 bytes (bulk-copy granule)");

// Packed env state (reference EnvState, gridworld.py:12-18): pos | exists_mask << 8 | time << 16.
// obj_poss is constant for tabular levels and early_term is always false in a carried state
// (a terminated env is auto-reset inside step), so neither is stored.
__host__ __device__ __forceinline__ int32_t pack_state(int pos, int exists, int time) {
    return pos | (exists << 8) | (time << 16);
}
__host__ __device__ __forceinline__ int st_pos(int32_t s) { return s & 0xff; }
__host__ __device__ __forceinline__ int st_exists(int32_t s) { return (s >> 8) & 0xff; }
__host__ __device__ __forceinline__ int st_time(int32_t s) { return (s >> 16) & 0xffff; }
// Packed observation: table row index | time << 16   (gridworld.py:184-205)
__host__ __device__ __forceinline__ int32_t pack_obs(int idx, int time) { return idx | (time << 16); }
__host__ __device__ __forceinline__ int ob_idx(int32_t o) { return o & 0xffff; }
__host__ __device__ __forceinline__ int ob_time(int32_t o) { return (o >> 16) & 0xffff; }

// ------------------------------------------------------------------------------------------------
// threefry2x32, 20 rounds (Random123; jax/_src/prng.py _threefry2x32_lowering)
struct Key { uint32_t a, b; };

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(x, x, r);
#else
    return (x << r) | (x >> (32 - r));
#endif
}

#define TF_ROUND(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }

__host__ __device__ __forceinline__ void threefry2x32(Key k, uint32_t& x0, uint32_t& x1) {
    const uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
    x0 += ks0; x1 += ks1;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks1; x1 += ks2 + 1u;
    TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
    x0 += ks2; x1 += ks0 + 2u;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks0; x1 += ks1 + 3u;
    TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
    x0 += ks1; x1 += ks2 + 4u;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks2; x1 += ks0 + 5u;
}

// element i of threefry_2x32(key, iota(n)) (jax 0.4.13, non-partitionable): the count vector is
// padded to even length with one 0, split into halves (x0 = first, x1 = second) and the two
// output halves are concatenated.
__host__ __device__ __forceinline__ uint32_t bits_elem(Key k, uint32_t n, uint32_t i) {
    const uint32_t half = (n + 1u) >> 1;
    const uint32_t blk = i < half ? i : i - half;
    uint32_t x0 = blk, x1 = half + blk;
    if (x1 >= n) x1 = 0u;                 // padding element
    threefry2x32(k, x0, x1);
    return i < half ? x0 : x1;
}

// jax.random.split(key, 2)
__host__ __device__ __forceinline__ void split2(Key k, Key& first, Key& second) {
    uint32_t a0 = 0u, b0 = 2u, a1 = 1u, b1 = 3u;
    threefry2x32(k, a0, b0);
    threefry2x32(k, a1, b1);
    first.a = a0; first.b = a1; second.a = b0; second.b = b1;
}
// jax.random.split(key, 2)[0] only (second key is dead)
__host__ __device__ __forceinline__ Key split2_first(Key k) {
    uint32_t a0 = 0u, b0 = 2u, a1 = 1u, b1 = 3u;
    threefry2x32(k, a0, b0);
    threefry2x32(k, a1, b1);
    Key r; r.a = a0; r.b = a1; return r;
}
// jax.random.split(key, n)[j]
__host__ __device__ __forceinline__ Key split_n(Key k, uint32_t n, uint32_t j) {
    Key r;
    r.a = bits_elem(k, 2u * n, 2u * j);
    r.b = bits_elem(k, 2u * n, 2u * j + 1u);
    return r;
}
// jax.random.split(key, 3) -> keys 0 and 1 (key 2 is dead on the tabular path)
__host__ __device__ __forceinline__ void split3_first_two(Key k, Key& k0, Key& k1) {
    uint32_t a0 = 0u, b0 = 3u, a1 = 1u, b1 = 4u, a2 = 2u, b2 = 5u;
    threefry2x32(k, a0, b0);
    threefry2x32(k, a1, b1);
    threefry2x32(k, a2, b2);
    k0.a = a0; k0.b = a1; k1.a = a2; k1.b = b0;
}

// jax _uniform: mantissa trick, [0, 1)
__host__ __device__ __forceinline__ float bits_to_unit(uint32_t bits) {
    const uint32_t fb = (bits >> 9) | 0x3F800000u;
#ifdef __CUDA_ARCH__
    return __fsub_rn(__uint_as_float(fb), 1.0f);
#else
    float f; memcpy(&f, &fb, 4); return f - 1.0f;
#endif
}
// jax.random.uniform(key, ()) : bits = threefry(key, [0] padded to [0,0]).x0
__host__ __device__ __forceinline__ float uniform_scalar(Key k) {
    uint32_t x0 = 0u, x1 = 0u;
    threefry2x32(k, x0, x1);
    return bits_to_unit(x0);
}

// ------------------------------------------------------------------------------------------------
// exp(x), x <= 0, individually rounded ops only (mirrors oracle/rollout.py::exp_portable)
__device__ __forceinline__ float exp_portable(float x) {
    x = fmaxf(x, -80.0f);
    const float n = rintf(__fmul_rn(x, 1.4426950408889634f));
    float r = __fsub_rn(x, __fmul_rn(n, 0.693359375f));
    r = __fsub_rn(r, __fmul_rn(n, -2.12194440e-4f));
    float p = 1.9841270e-4f;
    p = __fadd_rn(__fmul_rn(p, r), 1.3888889e-3f);
    p = __fadd_rn(__fmul_rn(p, r), 8.3333338e-3f);
    p = __fadd_rn(__fmul_rn(p, r), 4.1666668e-2f);
    p = __fadd_rn(__fmul_rn(p, r), 1.6666667e-1f);
    p = __fadd_rn(__fmul_rn(p, r), 0.5f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f);
    return __int_as_float(__float_as_int(p) + (__float2int_rn(n) << 23));
}

// softmax over C logits with the portable exp and a left-to-right sum (oracle: softmax_portable)
template <int C>
__device__ __forceinline__ void softmax_portable(const float (&z)[C], float (&p)[C]) {
    float m = z[0];
#pragma unroll
    for (int j = 1; j < C; ++j) m = fmaxf(m, z[j]);
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < C; ++j) { p[j] = exp_portable(__fsub_rn(z[j], m)); s = __fadd_rn(s, p[j]); }
#pragma unroll
    for (int j = 0; j < C; ++j) p[j] = __fdiv_rn(p[j], s);
}

// obs @ W for the one-hot-plus-time observation (models/agent.py:13-17 with actor_net=()):
// logits[j] = W[idx][j] + (time * 0.001) * W[D-1][j]
template <int C, int STRIDE>
__device__ __forceinline__ void tab_logits(const float* __restrict__ table, int D, int idx, int time,
                                           float (&z)[C]) {
    const float tf = __fmul_rn(__int2float_rn(time), 0.001f);
    const float* row = table + (size_t)idx * STRIDE;
    const float* last = table + (size_t)(D - 1) * STRIDE;
    if constexpr (STRIDE % 4 == 0) {
        float r[STRIDE], l[STRIDE];
#pragma unroll
        for (int q = 0; q < STRIDE / 4; ++q) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(row) + q);
            const float4 b = __ldg(reinterpret_cast<const float4*>(last) + q);
            r[4 * q] = a.x; r[4 * q + 1] = a.y; r[4 * q + 2] = a.z; r[4 * q + 3] = a.w;
            l[4 * q] = b.x; l[4 * q + 1] = b.y; l[4 * q + 2] = b.z; l[4 * q + 3] = b.w;
        }
#pragma unroll
        for (int j = 0; j < C; ++j) z[j] = __fadd_rn(r[j], __fmul_rn(tf, l[j]));
    } else {
#pragma unroll
        for (int j = 0; j < C; ++j) z[j] = __fadd_rn(__ldg(row + j), __fmul_rn(tf, __ldg(last + j)));
    }
}

// ------------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk copy (TMA) helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#if defined(TOUED_FUZZ)
// diagnostic build (csrc/build.py::VARIANTS["fuzz"]): one in eight barrier waits first sleeps for a pseudo-random
// time of up to 16 us, which shuffles the relative progress of the warp roles of every mbarrier-synchronised kernel.
// A correctly synchronised kernel gives bit-identical results under any such schedule.
__device__ __forceinline__ void fuzz_delay() {
    unsigned c = (unsigned)clock() ^ ((threadIdx.x >> 5) * 2654435761u) ^ (blockIdx.x * 40503u);
    c ^= c >> 13; c *= 0x5bd1e995u; c ^= c >> 15;
    if ((c & 7u) == 0) __nanosleep((c >> 8) & 0x3FFFu);
}
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if defined(TOUED_FUZZ)
    fuzz_delay();
#endif
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// host-side error plumbing (api.cu owns the storage)
void toued_set_error(const char* fmt, ...);
#define TOUED_CHECK(cond, ...) do { if (!(cond)) { toued_set_error(__VA_ARGS__); return 1; } } while (0)
#define TOUED_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    toued_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define TOUED_LAUNCH_CHECK() TOUED_CUDA(cudaGetLastError())
