// Device level generator: reference environments/gridworld/configs.py:12-53 `reset_env_params` (+ :56-57 reset_lifetime
// and environments/environments.py:23-38), one CTA per level, writing the packed LevelRec the rollout kernels read.
//
// The reference draws a level on every call of the level sampler for EVERY agent (level_sampler.py:152-153) and, with
// prioritised level replay, `minimum_new` = num_agents fresh buffer levels every meta-step (:331-353).  On the host
// (numpy + threefry) that was the largest cost of a GROOVE meta-step; here it is a 10 us kernel and
// the levels never leave the device.
//
// Bit-exact with oracle/configs.py (tests/test_16_levelgen_gpu.py) under the RNG / float contract of DESIGN.md section 2:
//   key tree    p_rng, l_rng = split(rng);  [distribution modes: rng, k = split(rng); sub-mode = randint(k, 0, n_modes)]
//               then one `rng, k = split(rng)` per parameter in the reference's order; callables that go through
//               `_sample_param` take one more split (configs.py:83-88);
//   uniform     mantissa trick, max(lo, u * (hi - lo) + lo) with individually rounded f32 operations;
//   log-uniform exp_portable(uniform(log lo, log hi)), integers by round-half-even;
//   randint     jax 0.4.13 `_randint` (two draws, multiplier (2^32 mod span)^2 mod span);
//   walls       permutation(G^2)[:n_walls] = one round of a stable sort by fresh 32-bit keys  -> rank counting;
//   positions   choice(G^2, (O + 1,), replace=False, p=valid) for a 0/1 mask = rank by decreasing 23-bit mantissa,
//               ties by index, inadmissible cells last                                         -> rank counting.
#include "common.cuh"
#include "../../include/toued.h"

struct GenParam  { int32_t kind, n; float lo, hi; float vals[8]; };      // 0 const vals[0..n) | 1 log-uniform (lo, hi = logs)
                                                                        // | 2 uniform | 3 uniform, first value in [0, hi)
struct GenScalar { int32_t kind, a, b; float lo, hi; };                  // 0 const a | 1 log-uniform int (logs) | 2 a + randint(b - a)
struct GenMode {
    GenParam reward, pterm, presp;
    GenScalar max_steps, n_objs, grid_size;
    int32_t wall_kind, n_walls;                                          // 0 fixed mask | 1 uniform without replacement
    uint32_t wallmask[8];
    int32_t obj_ids[8];
};
struct GenDesc {
    int32_t n_modes, O, T, G;
    GenScalar lifetime;
    GenMode modes[12];
};
static_assert(sizeof(GenParam) == 48 && sizeof(GenScalar) == 20 && sizeof(GenMode) == 276, "GenDesc layout is mirrored in Python");
static_assert(sizeof(GenDesc) == 36 + 12 * 276, "GenDesc layout is mirrored in Python");

__device__ __forceinline__ void next_key(Key& rng, Key& k) { Key a, b; split2(rng, a, b); rng = a; k = b; }
__device__ __forceinline__ Key second_of(Key k) { Key a, b; split2(k, a, b); return b; }

__device__ __forceinline__ float uniform_elem(Key k, uint32_t n, uint32_t i, float lo, float hi) {
    const float u = bits_to_unit(bits_elem(k, n, i));
    return fmaxf(lo, __fadd_rn(__fmul_rn(u, __fsub_rn(hi, lo)), lo));
}
// jax.random.randint(key, (), 0, span) for 0 < span < 2^31
__device__ __forceinline__ uint32_t randint0(Key key, uint32_t span) {
    Key k0, k1; split2(key, k0, k1);
    const uint64_t hi = bits_elem(k0, 1, 0), lo = bits_elem(k1, 1, 0);
    uint64_t m = (1ull << 32) % span;
    m = ((m * m) & 0xFFFFFFFFull) % span;
    uint64_t off = ((hi % span) * m) & 0xFFFFFFFFull;
    off = (off + (lo % span)) & 0xFFFFFFFFull;
    return (uint32_t)(off % span);
}
__device__ int32_t draw_scalar(Key k, const GenScalar& s, bool extra_split) {
    if (s.kind == 0) return s.a;
    if (extra_split) k = second_of(k);                                   // _sample_param: param(split(key)[1])
    if (s.kind == 1) return (int32_t)rintf(exp_portable(uniform_elem(k, 1, 0, s.lo, s.hi)));
    return s.a + (int32_t)randint0(k, (uint32_t)(s.b - s.a));
}
__device__ void draw_param(Key k, const GenParam& p, int T, float* out) {
    for (int i = 0; i < T; ++i) out[i] = 0.0f;                            // padded to the distribution's type count
    if (p.kind == 0) { for (int i = 0; i < p.n; ++i) out[i] = p.vals[i]; return; }
    if (p.kind == 1) { for (int i = 0; i < p.n; ++i) out[i] = exp_portable(uniform_elem(k, p.n, i, p.lo, p.hi)); return; }
    if (p.kind == 2) { for (int i = 0; i < p.n; ++i) out[i] = uniform_elem(k, p.n, i, p.lo, p.hi); return; }
    Key k0, k1; split2(k, k0, k1);                                        // uniform_first_pos (configs.py:98-107)
    out[0] = uniform_elem(k0, 1, 0, 0.0f, p.hi);
    for (int i = 1; i < p.n; ++i) out[i] = uniform_elem(k1, p.n - 1, i - 1, p.lo, p.hi);
}

__global__ void __launch_bounds__(256)
level_gen_kernel(const GenDesc* __restrict__ desc, const uint32_t* __restrict__ keys, const int32_t* __restrict__ buffer_ids,
                 LevelRec* __restrict__ out, int32_t* __restrict__ lifetimes_out, int n) {
    __shared__ GenDesc d;
    __shared__ uint32_t sbits[256];
    __shared__ long long srank[256];
    __shared__ uint32_t swalls[8];
    __shared__ int32_t s_mode, s_grid, s_pos[8];
    __shared__ Key s_kwall, s_kpos;
    __shared__ LevelRec rec;
    const int lvl = blockIdx.x, tid = threadIdx.x;
    if (lvl >= n) return;
    for (int i = tid; i < (int)(sizeof(GenDesc) / 4); i += 256) reinterpret_cast<uint32_t*>(&d)[i] = reinterpret_cast<const uint32_t*>(desc)[i];
    if (tid < 8) swalls[tid] = 0u;
    __syncthreads();
    const int O = d.O, T = d.T, G2 = d.G * d.G;
    if (tid == 0) {
        Key key; key.a = keys[2 * lvl]; key.b = keys[2 * lvl + 1];
        Key rng, lkey; split2(key, rng, lkey);                           // environments.py:30  p_rng, l_rng = split(rng)
        int mode = 0;
        if (d.n_modes > 1) { Key k; next_key(rng, k); mode = (int)randint0(k, (uint32_t)d.n_modes); }   // deviation Q4
        const GenMode& m = d.modes[mode];
        float rew[8], pt[8], pr[8];
        Key k;
        next_key(rng, k); draw_param(k, m.reward, T, rew);
        next_key(rng, k); draw_param(k, m.pterm, T, pt);
        next_key(rng, k); draw_param(k, m.presp, T, pr);
        next_key(rng, k); rec.max_steps = draw_scalar(k, m.max_steps, true);
        next_key(rng, k); rec.n_objs = draw_scalar(k, m.n_objs, true);
        next_key(rng, k); rec.grid_size = draw_scalar(k, m.grid_size, true);
        next_key(rng, k); s_kwall = second_of(k);                        // wall_idxs through _sample_param
        next_key(rng, k); s_kpos = k;
        rec.lifetime = draw_scalar(lkey, d.lifetime, false);
        rec.buffer_id = buffer_ids ? buffer_ids[lvl] : 0;
        rec._pad0[0] = rec._pad0[1] = 0;
        for (int i = 0; i < TOUED_MAX_OBJS; ++i) {
            float r = 0.f, a = 0.f, b = 0.f;
            if (i < O) {                                                  // jnp.take(params.obj_X, params.obj_ids): negative ids wrap
                int id = m.obj_ids[i];
                id = id < 0 ? id + T : id;
                id = min(max(id, 0), T - 1);
                r = rew[id]; a = pt[id]; b = pr[id];
            }
            rec.obj_reward[i] = r; rec.obj_p_term[i] = a; rec.obj_p_resp[i] = b;
            rec.obj_pos[i] = -1;
        }
        s_mode = mode; s_grid = rec.grid_size;
    }
    __syncthreads();
    const GenMode& m = d.modes[s_mode];
    // ---- walls ----
    if (m.wall_kind == 0) {
        if (tid < 8) swalls[tid] = m.wallmask[tid];
    } else {
        // permutation(G2)[:n_walls]: rounds = ceil(3 ln G2 / ln(2^32 - 1)) = 1 for G2 <= 1625; one round = stable sort by
        // bits(split(key)[1], (G2,)); cell i is a wall iff its rank is below n_walls
        const Key kb = second_of(s_kwall);
        if (tid < G2) sbits[tid] = bits_elem(kb, (uint32_t)G2, (uint32_t)tid);
        __syncthreads();
        if (tid < G2) {
            const uint32_t mine = sbits[tid];
            int rank = 0;
            for (int j = 0; j < G2; ++j) rank += (sbits[j] < mine || (sbits[j] == mine && j < tid)) ? 1 : 0;
            if (rank < m.n_walls) atomicOr(&swalls[tid >> 5], 1u << (tid & 31));
        }
    }
    __syncthreads();
    // ---- start + object positions: rank admissible cells by decreasing mantissa ----
    if (tid < G2) {
        const bool wall = (swalls[tid >> 5] >> (tid & 31)) & 1u;
        const bool valid = tid < s_grid * s_grid && !wall;
        const long long mant = (long long)(bits_elem(s_kpos, (uint32_t)G2, (uint32_t)tid) >> 9);
        srank[tid] = valid ? -mant : (1ll << 40);
    }
    __syncthreads();
    if (tid < G2) {
        const long long mine = srank[tid];
        int rank = 0;
        for (int j = 0; j < G2; ++j) rank += (srank[j] < mine || (srank[j] == mine && j < tid)) ? 1 : 0;
        if (rank <= O) s_pos[rank] = tid;
    }
    __syncthreads();
    if (tid == 0) {
        rec.start_pos = s_pos[0];
        for (int i = 0; i < O; ++i) rec.obj_pos[i] = s_pos[1 + i];
        for (int i = 0; i < 8; ++i) rec.walls[i] = swalls[i];
        if (lifetimes_out) lifetimes_out[lvl] = rec.lifetime;
    }
    __syncthreads();
    if (tid < (int)(sizeof(LevelRec) / 4)) reinterpret_cast<uint32_t*>(out + lvl)[tid] = reinterpret_cast<const uint32_t*>(&rec)[tid];
}

extern "C" int toued_generate_levels_desc_bytes(void) { return (int)sizeof(GenDesc); }

extern "C" int toued_generate_levels(const void* gen_desc, const uint32_t* keys, const int32_t* buffer_ids, void* levels_out,
                                     int32_t* lifetimes_out, int n_levels, void* stream) {
    TOUED_CHECK(gen_desc && keys && levels_out && n_levels > 0, "toued_generate_levels: bad arguments");
    level_gen_kernel<<<n_levels, 256, 0, (cudaStream_t)stream>>>((const GenDesc*)gen_desc, keys, buffer_ids,
                                                                 (LevelRec*)levels_out, lifetimes_out, n_levels);
    TOUED_LAUNCH_CHECK();
    return 0;
}
