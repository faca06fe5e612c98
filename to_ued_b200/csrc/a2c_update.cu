// A2C antagonist update for the tabular actor / value critic (one CTA per agent).
//
// Replaces (reference, JAX autodiff): agents/a2c.py:19-76 a2c_agent_train_step
//   critic: GAE targets (util/metrics.py:17-38), MSE on value[:-1] with stop-gradient targets, clip-SGD
//   actor : -log pi(a|s) * normalised advantage - entropy_coeff * entropy, clip-SGD, lifetime mask
// Quirk Q16 is reproduced: the value critic output keeps a trailing axis of size 1, so
// `-jnp.multiply(selected_log_probs, adv)` (a2c.py:60) is the [L, L] outer product per worker and its mean
// is -mean_t(log pi) * mean_t(adv)  (quirk = 0 gives the element-wise product).
// Closed-form sparse gradients + the deterministic segmented row reduction of segreduce.cuh.
#include "lpg_common.cuh"
#include "segreduce.cuh"
#include "../../include/toued.h"

constexpr int A2_C = 8;    // per-token record: 5 actor dlogits, 1 value dlogit, tf, pad

__global__ void __launch_bounds__(256)
a2c_update_kernel(const int32_t* __restrict__ obs, const uint8_t* __restrict__ action,
                  const float* __restrict__ reward, const uint8_t* __restrict__ done,
                  const uint16_t* __restrict__ sorted_tok, const float* __restrict__ actor_in,
                  const float* __restrict__ critic_in, float* actor_out, float* critic_out,
                  const LevelRec* __restrict__ levels, int32_t* __restrict__ step, float* __restrict__ scal,
                  int W, int L, int D, float lr_a, float lr_c, float max_norm, float gamma, float lmbda,
                  float ent_coeff, int quirk) {
    extern __shared__ __align__(16) float sm[];
    const int T = W * L;
    float* rec = sm;                        // [T][8]
    float* adv = rec + (size_t)T * A2_C;    // [T]
    float* abar = adv + T;                  // [W]
    float* runv = abar + W;                 // [T][6]
    float* scan = runv + (size_t)T * 6;     // [2][256][6]
    void* idxmem = scan + 2 * 256 * 6;
    __shared__ float red[32];
    __shared__ int iscan[512];
    __shared__ unsigned char sflags[512];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int32_t* ob = obs + (size_t)n * (L + 1) * W;
    const uint8_t* act = action + (size_t)n * T;
    const float* rw = reward + (size_t)n * T;
    const uint8_t* dn = done + (size_t)n * T;
    const float* a_in = actor_in + (size_t)n * D * 8;
    const float* c_in = critic_in + (size_t)n * D * 8;
    float* a_out = actor_out + (size_t)n * D * 8;
    float* c_out = critic_out + (size_t)n * D * 8;
    const float invT = 1.0f / (float)T;
    const SegIndex si = seg_index_build(idxmem, iscan, sorted_tok + (size_t)n * T, ob, T);
    for (int i = tid; i < D * 2; i += 256) {
        reinterpret_cast<float4*>(a_out)[i] = reinterpret_cast<const float4*>(a_in)[i];
        reinterpret_cast<float4*>(c_out)[i] = reinterpret_cast<const float4*>(c_in)[i];
    }
    // ---- critic: GAE per worker; d(mse)/dV_t = -2 adv_t / T  (targets are stop-gradient) ----
    float s1 = 0.f, s2 = 0.f;
    const float vlast = c_in[(size_t)(D - 1) * 8];
    for (int w = tid; w < W; w += 256) {
        float g = 0.f;
        const int32_t o1 = ob[L * W + w];
        float v1 = c_in[(size_t)ob_idx(o1) * 8] + 0.001f * (float)ob_time(o1) * vlast;
        for (int t = L - 1; t >= 0; --t) {
            const int32_t o0 = ob[t * W + w];
            const float v0 = c_in[(size_t)ob_idx(o0) * 8] + 0.001f * (float)ob_time(o0) * vlast;
            const float nd = dn[t * W + w] ? 0.0f : 1.0f;
            const float delta = rw[t * W + w] + gamma * v1 * nd - v0;
            g = delta + gamma * lmbda * nd * g;
            adv[t * W + w] = g;
            s1 += g; s2 = fmaf(g, g, s2);
            v1 = v0;
        }
    }
    s1 = block_sum(s1, red); s2 = block_sum(s2, red);
    const float mean = s1 * invT;
    float var = 0.f;
    for (int i = tid; i < T; i += 256) { const float d = adv[i] - mean; var = fmaf(d, d, var); }
    var = block_sum(var, red) * invT;
    const float inv_std = 1.0f / (sqrtf(var) + 1e-8f);
    __syncthreads();
    // value dlogit uses the raw advantage; then normalise in place
    for (int i = tid; i < T; i += 256) {
        rec[i * A2_C + 5] = -2.0f * invT * adv[i];
        adv[i] = (adv[i] - mean) * inv_std;
    }
    __syncthreads();
    for (int w = tid; w < W; w += 256) {
        float sa = 0.f;
        for (int t = 0; t < L; ++t) sa += adv[t * W + w];
        abar[w] = sa / (float)L;
    }
    __syncthreads();
    // ---- actor dlogits ----
    float last[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, aloss = 0.f;
    for (int tok = tid; tok < T; tok += 256) {
        const int32_t o = ob[tok];
        float z[5], p[5];
        tab_logits8<5>(a_in, D, o, z);
        softmax_c<5>(z, p);
        const int a = act[tok];
        float pa = p[0];
#pragma unroll
        for (int j = 1; j < 5; ++j) pa = (a == j) ? p[j] : pa;
        const float wgt = quirk ? abar[tok % W] : adv[tok];
        const float q = pa / (pa + 1e-8f);
        float dh[5], s = 0.f, ent = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float lp = logf(p[j] + 1e-8f);
            dh[j] = -(lp + 1.0f); s = fmaf(p[j], dh[j], s); ent -= (p[j] + 1e-8f) * lp;
        }
        float* r = rec + tok * A2_C;
#pragma unroll
        for (int j = 0; j < 5; ++j)
            r[j] = invT * (-wgt * q * ((a == j ? 1.0f : 0.0f) - p[j]) - ent_coeff * p[j] * (dh[j] - s));
        const float tf = 0.001f * (float)ob_time(o);
        r[6] = tf;
#pragma unroll
        for (int j = 0; j < 6; ++j) last[j] = fmaf(tf, r[j], last[j]);
        aloss += invT * (-wgt * logf(pa + 1e-8f) - ent_coeff * ent);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) last[j] = block_sum(last[j], red);
    aloss = block_sum(aloss, red);
    __syncthreads();
    seg_reduce<6, A2_C>(rec, si, T, runv, scan, sflags);
    float na = 0.f, nc = 0.f;
    for (int r = tid; r < si.nruns; r += 256) {
        const float* g = runv + r * 6;
#pragma unroll
        for (int j = 0; j < 5; ++j) na = fmaf(g[j], g[j], na);
        nc = fmaf(g[5], g[5], nc);
    }
    na = block_sum(na, red); nc = block_sum(nc, red);
#pragma unroll
    for (int j = 0; j < 5; ++j) na = fmaf(last[j], last[j], na);
    nc = fmaf(last[5], last[5], nc);
    const float gna = sqrtf(na), gnc = sqrtf(nc);
    const float sa_ = gna < max_norm ? 1.0f : max_norm / gna;
    const float sc_ = gnc < max_norm ? 1.0f : max_norm / gnc;
    const int old_step = step[n];
    const bool keep = (old_step + 1) <= levels[n].lifetime;            // a2c.py:70-75
    const float ua = keep ? lr_a * sa_ : 0.0f, uc = keep ? lr_c * sc_ : 0.0f;
    for (int r = tid; r < si.nruns; r += 256) {
        const int row = si.run_row[r];
        const float* g = runv + r * 6;
#pragma unroll
        for (int j = 0; j < 5; ++j) a_out[(size_t)row * 8 + j] = a_in[(size_t)row * 8 + j] - ua * g[j];
        c_out[(size_t)row * 8] = c_in[(size_t)row * 8] - uc * g[5];
    }
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) a_out[(size_t)(D - 1) * 8 + j] = a_in[(size_t)(D - 1) * 8 + j] - ua * last[j];
        c_out[(size_t)(D - 1) * 8] = c_in[(size_t)(D - 1) * 8] - uc * last[5];
        step[n] = keep ? old_step + 1 : old_step;
        scal[n * 4 + 0] = aloss;            // actor_loss
        scal[n * 4 + 1] = s2 * invT;        // critic_loss (target - value == adv)
        scal[n * 4 + 2] = gna; scal[n * 4 + 3] = gnc;
    }
}

extern "C" int toued_a2c_update(const int32_t* obs, const uint8_t* action, const float* reward, const uint8_t* done,
                                const uint16_t* sorted_tok, const float* actor_in, const float* critic_in,
                                float* actor_out, float* critic_out, const void* levels, int32_t* step,
                                float* scalars, int n_agents, int n_workers, int rollout_len, int obs_dim,
                                float lr_actor, float lr_critic, float max_grad_norm, float gamma, float gae_lambda,
                                float entropy_coeff, int outer_product_quirk, void* stream) {
    const int T = n_workers * rollout_len;
    TOUED_CHECK(n_agents > 0 && T > 0, "toued_a2c_update: empty problem");
    TOUED_CHECK(actor_in != actor_out && critic_in != critic_out, "toued_a2c_update: in-place update not supported");
    const size_t smem = sizeof(float) * ((size_t)T * A2_C + T + n_workers + (size_t)T * 6 + 2 * 256 * 6) + seg_index_bytes(T);
    TOUED_CHECK(smem <= 200 * 1024, "toued_a2c_update: W*L=%d too large for shared memory", T);
    TOUED_CUDA(cudaFuncSetAttribute(a2c_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a2c_update_kernel<<<n_agents, 256, smem, (cudaStream_t)stream>>>(
        obs, action, reward, done, sorted_tok, actor_in, critic_in, actor_out, critic_out, (const LevelRec*)levels,
        step, scalars, n_workers, rollout_len, obs_dim, lr_actor, lr_critic, max_grad_norm, gamma, gae_lambda,
        entropy_coeff, outer_product_quirk);
    TOUED_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// agents/a2c.py:79-125 as ONE call: the whole lifetime of an A2C agent (num_updates x [rollout -> token sort ->
// A2C update], tables ping-ponging between two buffers) is enqueued from native code.  The antagonist of the
// algorithmic-regret score trains for up to 2500 updates (mazes); launched from Python the loop is host-bound
// (~90 us per update in ctypes + tensor bookkeeping against ~25 us of kernels).
__global__ void a2c_accumulate_kernel(const float* __restrict__ scal, float* __restrict__ sums, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { sums[2 * i] += scal[4 * i]; sums[2 * i + 1] += scal[4 * i + 1]; }
}

extern "C" int toued_a2c_train(const void* levels, const uint32_t* keys, float* actor0, float* actor1, float* critic0,
                               float* critic1, int32_t* state, int32_t* obs, uint8_t* action, float* reward,
                               uint8_t* done, uint16_t* sorted_tok, int32_t* step, float* scalars, float* loss_sums,
                               int num_updates, int n_agents, int n_workers, int rollout_len, int obs_dim,
                               int max_grid_size, int max_n_objs, float lr_actor, float lr_critic, float max_grad_norm,
                               float gamma, float gae_lambda, float entropy_coeff, int outer_product_quirk, void* stream) {
    TOUED_CHECK(num_updates >= 0 && n_agents > 0, "toued_a2c_train: bad arguments");
    float* a[2] = {actor0, actor1};
    float* c[2] = {critic0, critic1};
    for (int k = 0; k < num_updates; ++k) {
        const int i = k & 1, o = i ^ 1;
        int rc = toued_rollout(levels, keys + (size_t)k * n_agents * 2, a[i], nullptr, state, obs, action, reward, done, nullptr,
                               n_agents, n_workers, rollout_len, obs_dim, max_grid_size, max_n_objs, 0, stream);
        if (rc) return rc;
        rc = toued_sort_tokens(obs, sorted_tok, n_agents, n_workers, rollout_len, stream);
        if (rc) return rc;
        rc = toued_a2c_update(obs, action, reward, done, sorted_tok, a[i], c[i], a[o], c[o], levels, step, scalars, n_agents,
                              n_workers, rollout_len, obs_dim, lr_actor, lr_critic, max_grad_norm, gamma, gae_lambda,
                              entropy_coeff, outer_product_quirk, stream);
        if (rc) return rc;
        a2c_accumulate_kernel<<<(n_agents + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scalars, loss_sums, n_agents);
        TOUED_LAUNCH_CHECK();
    }
    return 0;
}
