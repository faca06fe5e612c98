// LPG-driven agent update for the tabular actor / critic (one CTA per agent).
//
// Replaces (reference, JAX autodiff over dense one-hot matmuls):
//   agents/lpg_agent.py:31-85     lpg_agent_train_step   (loss, grads wrt actor & critic, SGD, lifetime mask)
//   agents/lpg_agent.py:119-120   batch_rollout_entropy of the UPDATED nets on the rollout's obs
//   models/optim.py:6-11          SGD = clip_by_global_norm -> scale(lr) -> scale(-1)      [optax 0.1.5]
//   util/metrics.py:5-14          entropy, kl_divergence
//
// With actor_net = critic_net = () the nets are D x C tables and a token touches two rows (its
// observation row and the time row D-1), so the gradient is closed-form and sparse:
//   actor : dlogit_j = (1/T) * pi_hat * p_a/(p_a+1e-8) * (delta_aj - p_j)
//   critic: dlogit_i = (alpha/T) * y_i * (m_i - sum_j y_j m_j),  m = log(y+e) - log(y_hat+e) + y/(y+e)
// Row sums are segmented sums over the agent's row-sorted token list (toued_sort_tokens), computed by
// the fixed-tree block reduction of segreduce.cuh: bitwise deterministic, no atomics.
#include "lpg_common.cuh"
#include "segreduce.cuh"
#include "../../include/toued.h"

constexpr int AU_C = 13;   // per-token record: 5 actor dlogits, 8 critic dlogits

__global__ void __launch_bounds__(256, 2)
agent_update_kernel(const int32_t* __restrict__ obs, const uint8_t* __restrict__ action,
                    const uint16_t* __restrict__ sorted_tok, const float* __restrict__ pi_hat,
                    const float* __restrict__ y_hat, const float* __restrict__ actor_in,
                    const float* __restrict__ critic_in, float* actor_out, float* critic_out,
                    const LevelRec* __restrict__ levels, int32_t* __restrict__ step,
                    float* __restrict__ scal, float* run_scratch, int n_agents, int W, int L, int D,
                    float lr_a, float lr_c, float max_norm, float alpha, int tables_precopied) {
    extern __shared__ __align__(16) float smc[];          // [T][AU_C] records | scan | index   (106 KB at T = 1280: two CTAs per SM)
    __shared__ float red[32];
    __shared__ int iscan[512];
    __shared__ unsigned char sflags[512];
    const int n = blockIdx.x, tid = threadIdx.x, T = W * L, R = n_agents * W;
    // run sums [min(T, D)][13] (a run is a distinct table row: at most D of them): L2-resident global scratch
    float* runv = run_scratch + (size_t)n * min(T, D) * 13;
    float* scan = smc + (size_t)T * AU_C;
    void* idxmem = scan + 2 * 256 * 13;
    const int32_t* ob = obs + (size_t)n * (L + 1) * W;
    const uint8_t* act = action + (size_t)n * T;
    const float* a_in = actor_in + (size_t)n * D * 8;
    const float* c_in = critic_in + (size_t)n * D * 8;
    float* a_out = actor_out + (size_t)n * D * 8;
    float* c_out = critic_out + (size_t)n * D * 8;
    const float invT = 1.0f / (float)T;
    const SegIndex si = seg_index_build(idxmem, iscan, sorted_tok + (size_t)n * T, ob, T);

    // dense copy theta_k -> theta_{k+1} (only touched rows change below).  The production step makes this copy on a side
    // stream while the LPG forward runs (tables_precopied: it is ~100 MB of pure traffic per launch, a third of this
    // kernel's time, and depends on nothing but theta_k)
    if (!tables_precopied)
        for (int i = tid; i < D * 2; i += 256) {
            reinterpret_cast<float4*>(a_out)[i] = reinterpret_cast<const float4*>(a_in)[i];
            reinterpret_cast<float4*>(c_out)[i] = reinterpret_cast<const float4*>(c_in)[i];
        }

    // ---- phase 0: per-token dlogits + metric partials -------------------------------------
    float m_kl = 0.f, m_pi2 = 0.f, m_y2 = 0.f;
    float last[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) last[j] = 0.f;
    for (int tok = tid; tok < T; tok += 256) {
        const int t = tok / W, w = tok - t * W;
        const size_t li = (size_t)t * R + (size_t)n * W + w;
        const int32_t o = ob[tok];
        float z[5], p[5], zy[8], y[8];
        tab_logits8<5>(a_in, D, o, z);
        softmax_c<5>(z, p);
        tab_logits8<8>(c_in, D, o, zy);
        softmax_c<8>(zy, y);
        const int a = act[tok];
        float pa = p[0];
#pragma unroll
        for (int j = 1; j < 5; ++j) pa = (a == j) ? p[j] : pa;
        const float ph = pi_hat[li];
        const float ca = invT * ph * (pa / (pa + 1e-8f));
        float* rec = smc + tok * AU_C;
#pragma unroll
        for (int j = 0; j < 5; ++j) rec[j] = ca * ((a == j ? 1.0f : 0.0f) - p[j]);
        float yh[8];
        { const float4* q = reinterpret_cast<const float4*>(y_hat + li * 8); const float4 q0 = q[0], q1 = q[1];
          yh[0] = q0.x; yh[1] = q0.y; yh[2] = q0.z; yh[3] = q0.w; yh[4] = q1.x; yh[5] = q1.y; yh[6] = q1.z; yh[7] = q1.w; }
        float mm[8], b = 0.f, kl = 0.f, y2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float ly = logf(y[i] + 1e-8f), lh = logf(yh[i] + 1e-8f);
            mm[i] = ly - lh + y[i] / (y[i] + 1e-8f);
            b = fmaf(y[i], mm[i], b);
            kl = fmaf(y[i], ly - lh, kl);
            y2 = fmaf(yh[i], yh[i], y2);
        }
        const float cc = alpha * invT;
#pragma unroll
        for (int i = 0; i < 8; ++i) rec[5 + i] = cc * y[i] * (mm[i] - b);
        const float tf = 0.001f * (float)ob_time(o);
#pragma unroll
        for (int j = 0; j < 13; ++j) last[j] = fmaf(tf, rec[j], last[j]);
        m_kl += kl; m_pi2 = fmaf(ph, ph, m_pi2); m_y2 += y2;
    }
    // time-row gradient (row D-1) + metric sums (deterministic tree)
#pragma unroll
    for (int j = 0; j < 13; ++j) last[j] = block_sum(last[j], red);
    m_kl = block_sum(m_kl, red); m_pi2 = block_sum(m_pi2, red); m_y2 = block_sum(m_y2, red);
    __syncthreads();

    // ---- phase 1: row gradients (segmented sums) and their squared norms --------------------------
    seg_reduce<13, AU_C>(smc, si, T, runv, scan, sflags);
    float na = 0.f, nc = 0.f;
    for (int r = tid; r < si.nruns; r += 256) {
        const float* g = runv + r * 13;
#pragma unroll
        for (int j = 0; j < 5; ++j) na = fmaf(g[j], g[j], na);
#pragma unroll
        for (int j = 5; j < 13; ++j) nc = fmaf(g[j], g[j], nc);
    }
    na = block_sum(na, red); nc = block_sum(nc, red);
#pragma unroll
    for (int j = 0; j < 5; ++j) na = fmaf(last[j], last[j], na);
#pragma unroll
    for (int j = 5; j < 13; ++j) nc = fmaf(last[j], last[j], nc);
    const float gna = sqrtf(na), gnc = sqrtf(nc);
    // optax clip_by_global_norm: g if ||g|| < c else g / ||g|| * c
    const float sa = gna < max_norm ? 1.0f : max_norm / gna;
    const float sc = gnc < max_norm ? 1.0f : max_norm / gnc;
    const int old_step = step[n];
    const bool keep = (old_step + 1) <= levels[n].lifetime;          // lpg_agent.py:78-82
    const float ua = keep ? lr_a * sa : 0.0f, uc = keep ? lr_c * sc : 0.0f;

    // ---- phase 2: SGD on the touched rows ---------------------------------------------------
    for (int r = tid; r < si.nruns; r += 256) {
        const int row = si.run_row[r];
        const float* g = runv + r * 13;
#pragma unroll
        for (int j = 0; j < 5; ++j) a_out[(size_t)row * 8 + j] = a_in[(size_t)row * 8 + j] - ua * g[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) c_out[(size_t)row * 8 + j] = c_in[(size_t)row * 8 + j] - uc * g[5 + j];
    }
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) a_out[(size_t)(D - 1) * 8 + j] = a_in[(size_t)(D - 1) * 8 + j] - ua * last[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) c_out[(size_t)(D - 1) * 8 + j] = c_in[(size_t)(D - 1) * 8 + j] - uc * last[5 + j];
        step[n] = keep ? old_step + 1 : old_step;
    }
    __syncthreads();     // block-scope visibility of the updated rows

    // ---- phase 3: entropies of the updated nets on this rollout's observations ----------------
    float ea = 0.f, ec = 0.f;
    for (int tok = tid; tok < T; tok += 256) {
        const int32_t o = ob[tok];
        float z[5], p[5], zy[8], y[8];
        tab_logits8<5>(a_out, D, o, z);
        softmax_c<5>(z, p);
        tab_logits8<8>(c_out, D, o, zy);
        softmax_c<8>(zy, y);
#pragma unroll
        for (int j = 0; j < 5; ++j) { const float q = p[j] + 1e-8f; ea -= q * logf(q); }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float q = y[j] + 1e-8f; ec -= q * logf(q); }
    }
    ea = block_sum(ea, red); ec = block_sum(ec, red);
    if (tid == 0) {
        float* s = scal + (size_t)n * 8;
        s[0] = gna; s[1] = gnc; s[2] = keep ? 1.0f : 0.0f;
        s[3] = m_kl * invT;      // critic_loss
        s[4] = m_pi2 * invT;     // policy_l2 (pi_l2)
        s[5] = m_y2 * invT;      // critic_l2 (y_l2)
        s[6] = ea * invT;        // policy_entropy
        s[7] = ec * invT;        // critic_entropy
    }
}

// 108 KB at T = 1280, D = 101: two CTAs per SM, so the 256 agents of a launch are resident at once
static size_t agent_smem_bytes(int T, int D) {
    return sizeof(float) * ((size_t)T * AU_C + 2 * 256 * 13) + seg_index_bytes(T);
}

extern "C" int toued_agent_update(const int32_t* obs, const uint8_t* action, const uint16_t* sorted_tok,
                                  const float* pi_hat, const float* y_hat, const float* actor_in,
                                  const float* critic_in, float* actor_out, float* critic_out,
                                  const void* levels, int32_t* step, float* scalars, int n_agents,
                                  int n_workers, int rollout_len, int obs_dim, float lr_actor,
                                  float lr_critic, float max_grad_norm, float agent_target_coeff, int tables_precopied,
                                  float* run_scratch, void* stream) {
    const int T = n_workers * rollout_len;
    const size_t smem = agent_smem_bytes(T, obs_dim);
    TOUED_CHECK(n_agents > 0 && T > 0, "toued_agent_update: empty problem");
    TOUED_CHECK(smem <= 200 * 1024, "toued_agent_update: W*L=%d too large for shared memory", T);
    TOUED_CHECK(actor_in != actor_out && critic_in != critic_out, "toued_agent_update: in-place update not supported");
    TOUED_CUDA(cudaFuncSetAttribute(agent_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TOUED_CUDA(cudaFuncSetAttribute(agent_update_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    TOUED_CHECK(run_scratch != nullptr, "toued_agent_update: run_scratch (toued_agent_scratch_floats) is required");
    cudaStream_t st = (cudaStream_t)stream;
    agent_update_kernel<<<n_agents, 256, smem, st>>>(
        obs, action, sorted_tok, pi_hat, y_hat, actor_in, critic_in, actor_out, critic_out,
        (const LevelRec*)levels, step, scalars, run_scratch, n_agents, n_workers, rollout_len, obs_dim,
        lr_actor, lr_critic, max_grad_norm, agent_target_coeff, tables_precopied);
    TOUED_LAUNCH_CHECK();
    return 0;
}

// floats of caller-provided scratch for one toued_agent_update / toued_agent_backward launch: the per-row run sums
// [min(W*L, D)][13] of every agent (written and read inside the launch only; L2-resident)
extern "C" int toued_agent_scratch_floats(int n_agents, int n_workers, int rollout_len, int obs_dim) {
    const long long T = (long long)n_workers * rollout_len;
    const long long n = 13ll * (T < obs_dim ? T : obs_dim) * n_agents;
    return n < (1ll << 31) ? (int)n : -1;
}
