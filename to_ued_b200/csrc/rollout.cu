// Fused gridworld rollout: reset-select / step / policy forward / action sampling / trajectory
// write, one thread per environment, the whole L-step scan inside one kernel.
//
// Replaces (reference, JAX):
//   environments/rollout.py:45-102      RolloutWrapper.batch_rollout / single_rollout
//   environments/gridworld/gridworld.py:72-211  step_env / reset_env / get_obs / _get_next_pos
//   gymnax==0.0.6 Environment.step      key split + auto-reset select            [3P-recall]
//   models/agent.py:7-17                Actor with actor_net=() (row gather + softmax)
//
// Layout: thread g -> (agent = g / W, worker = g % W).  A block owns APB consecutive agents and
// stages their LevelRecs (192 B each) in shared memory with one cp.async.bulk + mbarrier.  Env
// state (pos, exists mask, time), the RNG key chain and the running first-episode return live in
// registers for all L steps.  The actor table row gather goes through the read-only path
// (two 32 B-aligned float4 loads for the row, two for the time row; tables are L2-resident).
// Trajectories are written [agent][t][worker]: consecutive threads write consecutive addresses.
//
// This file is compiled with -fmad=false: the sampling path must round exactly like the oracle.
#include "gridworld.cuh"
#include "../../include/toued.h"

template <int O>
__global__ void __launch_bounds__(128)
rollout_kernel(const LevelRec* __restrict__ levels, const uint32_t* __restrict__ keys,
               const float* __restrict__ actor, const uint8_t* __restrict__ forced_actions,
               int32_t* __restrict__ state, int32_t* __restrict__ obs, uint8_t* __restrict__ action,
               float* __restrict__ reward, uint8_t* __restrict__ done, float* __restrict__ ep_return,
               int n_agents, int W, int L, int D, int G2, int apb, int reset_first) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelRec* lv = reinterpret_cast<LevelRec*>(smem_raw);
    __shared__ __align__(8) uint64_t bar;

    const int agent0 = blockIdx.x * apb;
    const int n_here = min(apb, n_agents - agent0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(n_here * sizeof(LevelRec));
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(lv, levels + agent0, bytes, &bar);
    }
    mbar_wait(&bar, 0);

    const int la = threadIdx.x / W;             // agent within block
    const int w = threadIdx.x - la * W;
    const int agent = agent0 + la;
    if (la >= n_here) return;
    const LevelRec& lev = lv[la];
    const float* table = actor + (size_t)agent * D * TOUED_ACT_PAD;

    // per-worker key: jax.random.split(rng, W)[w]                      rollout.py:49
    Key rng;
    {
        Key ak; ak.a = keys[2 * agent]; ak.b = keys[2 * agent + 1];
        rng = split_n(ak, (uint32_t)W, (uint32_t)w);
    }

    EnvRegs s;
    if (reset_first) {                          // batch_reset (rollout.py:38-42), tabular: deterministic
        s = env_reset(lev);
    } else {
        const int32_t ps = state[(size_t)agent * W + w];
        s.pos = st_pos(ps); s.exists = st_exists(ps); s.time = st_time(ps);
    }
    float cum = 0.0f, valid = 1.0f;
    const size_t tok0 = (size_t)agent * L * W + w;          // [agent][t][w]
    const size_t ob0 = (size_t)agent * (L + 1) * W + w;     // [agent][t (L+1)][w]
    const bool write = obs != nullptr;
    if (write) obs[ob0] = pack_obs(obs_row(s, G2), s.time);

    for (int t = 0; t < L; ++t) {
        // ---- RNG for this step (independent of the env dynamics) -------- rollout.py:61-66
        Key k_act, k_env;
        split2(rng, rng, k_act);
        split2(rng, rng, k_env);
        const float u_act = uniform_scalar(k_act);
        const StepRand<O> rnd = env_step_rand<O>(k_env);

        // ---- policy forward + action sampling ---------------------------- rollout.py:62-63
        int a;
        {
            float z[TOUED_NUM_ACTIONS], p[TOUED_NUM_ACTIONS];
            tab_logits<TOUED_NUM_ACTIONS, TOUED_ACT_PAD>(table, D, obs_row(s, G2), s.time, z);
            softmax_portable<TOUED_NUM_ACTIONS>(z, p);
            float c[TOUED_NUM_ACTIONS];
            c[0] = p[0];
#pragma unroll
            for (int j = 1; j < TOUED_NUM_ACTIONS; ++j) c[j] = __fadd_rn(c[j - 1], p[j]);
            const float r = __fmul_rn(c[TOUED_NUM_ACTIONS - 1], __fsub_rn(1.0f, u_act));
            a = 0;
#pragma unroll
            for (int j = 0; j < TOUED_NUM_ACTIONS; ++j) a += (c[j] < r) ? 1 : 0;
        }
        if (forced_actions) a = forced_actions[tok0 + (size_t)t * W];

        // ---- env.step (auto-resetting) ------------------------------------ rollout.py:65
        bool dn;
        const float rew = env_step_apply<O>(lev, rnd, a, s, dn);

        cum = __fadd_rn(cum, __fmul_rn(rew, valid));        // rollout.py:68-69
        valid = __fmul_rn(valid, dn ? 0.0f : 1.0f);

        if (write) {
            const size_t k = tok0 + (size_t)t * W;
            action[k] = (uint8_t)a;
            reward[k] = rew;
            done[k] = dn ? 1 : 0;
            obs[ob0 + (size_t)(t + 1) * W] = pack_obs(obs_row(s, G2), s.time);
        }
    }
    if (state) state[(size_t)agent * W + w] = pack_state(s.pos, s.exists, s.time);
    if (ep_return) ep_return[(size_t)agent * W + w] = cum;
}

extern "C" int toued_rollout(const void* levels, const uint32_t* keys, const float* actor,
                             const uint8_t* forced_actions, int32_t* state, int32_t* obs,
                             uint8_t* action, float* reward, uint8_t* done, float* ep_return,
                             int n_agents, int n_workers, int rollout_len, int obs_dim,
                             int max_grid_size, int max_n_objs, int reset_first, void* stream) {
    TOUED_CHECK(n_agents > 0 && n_workers > 0 && rollout_len > 0, "toued_rollout: empty problem");
    TOUED_CHECK(n_workers <= 128 && (128 % n_workers) == 0,
                "toued_rollout: n_workers=%d must divide 128", n_workers);
    TOUED_CHECK(max_n_objs >= 1 && max_n_objs <= 5, "toued_rollout: max_n_objs=%d not in 1..5", max_n_objs);
    TOUED_CHECK(max_grid_size * max_grid_size <= 255, "toued_rollout: max_grid_size too large");
    TOUED_CHECK(obs_dim == max_grid_size * max_grid_size * (1 << max_n_objs) + 1,
                "toued_rollout: obs_dim=%d inconsistent with grid/objs", obs_dim);
    TOUED_CHECK(rollout_len < 65536, "toued_rollout: rollout_len too large");
    TOUED_CHECK(obs == nullptr || (action && reward && done), "toued_rollout: partial trajectory buffers");
    const int apb = 128 / n_workers;
    const int blocks = (n_agents + apb - 1) / apb;
    const size_t smem = (size_t)apb * sizeof(LevelRec);
    const int G2 = max_grid_size * max_grid_size;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(OO) rollout_kernel<OO><<<blocks, 128, smem, st>>>( \
        (const LevelRec*)levels, keys, actor, forced_actions, state, obs, action, reward, done, ep_return, \
        n_agents, n_workers, rollout_len, obs_dim, G2, apb, reset_first)
    switch (max_n_objs) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(5); break;
    }
#undef LAUNCH
    TOUED_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// gymnax-style single transitions (environments/gridworld/gridworld.py + gymnax Environment.step):
// one thread per env, per-env keys given explicitly.  API completeness; not on the hot path.
template <int O>
__global__ void env_step_kernel(const LevelRec* __restrict__ levels, const uint32_t* __restrict__ keys,
                                const int32_t* __restrict__ actions, int32_t* __restrict__ state,
                                int32_t* __restrict__ obs, float* __restrict__ reward,
                                uint8_t* __restrict__ done, int n_envs, int W, int G2) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_envs) return;
    const LevelRec lev = levels[g / W];
    Key k; k.a = keys[2 * g]; k.b = keys[2 * g + 1];
    const int32_t ps = state[g];
    EnvRegs s; s.pos = st_pos(ps); s.exists = st_exists(ps); s.time = st_time(ps);
    const StepRand<O> rnd = env_step_rand<O>(k);
    bool dn;
    const float rew = env_step_apply<O>(lev, rnd, actions[g], s, dn);
    state[g] = pack_state(s.pos, s.exists, s.time);
    obs[g] = pack_obs(obs_row(s, G2), s.time);
    reward[g] = rew;
    done[g] = dn ? 1 : 0;
}

__global__ void env_reset_kernel(const LevelRec* __restrict__ levels, int32_t* __restrict__ state,
                                 int32_t* __restrict__ obs, int n_envs, int W, int G2) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_envs) return;
    const LevelRec lev = levels[g / W];
    const EnvRegs s = env_reset(lev);
    state[g] = pack_state(s.pos, s.exists, s.time);
    if (obs) obs[g] = pack_obs(obs_row(s, G2), s.time);
}

extern "C" int toued_env_step(const void* levels, const uint32_t* keys, const int32_t* actions,
                              int32_t* state, int32_t* obs, float* reward, uint8_t* done,
                              int n_agents, int n_workers, int max_grid_size, int max_n_objs, void* stream) {
    TOUED_CHECK(n_agents > 0 && n_workers > 0, "toued_env_step: empty problem");
    TOUED_CHECK(max_n_objs >= 1 && max_n_objs <= 5, "toued_env_step: max_n_objs=%d not in 1..5", max_n_objs);
    const int n = n_agents * n_workers, G2 = max_grid_size * max_grid_size;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(OO) env_step_kernel<OO><<<(n + 127) / 128, 128, 0, st>>>((const LevelRec*)levels, keys, actions, \
        state, obs, reward, done, n, n_workers, G2)
    switch (max_n_objs) {
        case 1: LAUNCH(1); break; case 2: LAUNCH(2); break; case 3: LAUNCH(3); break;
        case 4: LAUNCH(4); break; default: LAUNCH(5); break;
    }
#undef LAUNCH
    TOUED_LAUNCH_CHECK();
    return 0;
}

extern "C" int toued_env_reset(const void* levels, int32_t* state, int32_t* obs, int n_agents,
                               int n_workers, int max_grid_size, void* stream) {
    TOUED_CHECK(n_agents > 0 && n_workers > 0, "toued_env_reset: empty problem");
    const int n = n_agents * n_workers;
    env_reset_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        (const LevelRec*)levels, state, obs, n, n_workers, max_grid_size * max_grid_size);
    TOUED_LAUNCH_CHECK();
    return 0;
}
