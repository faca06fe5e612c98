// Meta-loss and the adjoint (reverse-mode) recursion through the K agent updates, one CTA per agent.
//
// Replaces what the reference obtains from jax.grad over its unrolled inner loop:
//   meta/train.py:60-100           value "update" (Q2: parameters never change), GAE advantage,
//                                  normalisation, LPG loss (Q16: outer-product broadcast), regularisers
//   agents/agents.py:109-116       compute_advantage ;  util/metrics.py:17-38 gae
//   agents/lpg_agent.py:60-82      backward through grad / clip_by_global_norm / SGD / lifetime mask
//
// Notation.  theta_k, phi_k: actor / critic tables before update k.  lam, mu: adjoints dL/dtheta,
// dL/dphi (dense [D][8] tables in global memory, touched row-wise).  For update k
//     theta_{k+1} = theta_k - keep * lr * C(g_k),  g_k = grad_theta mean(log(pi_theta(a|s)+1e-8) * pi_hat_k)
// so with w = -lr * J_C(g_k)^T lam_{k+1}:
//     d L / d pi_hat_k[tok] = <w, grad_theta log(pi+1e-8)[tok]> / T        (-> LPG backward)
//     lam_k = lam_{k+1} + grad_theta <w, g_k(theta)>                        (Hessian-vector product)
// and likewise for the critic with the KL(y_t || y_hat) loss.  All closed form for softmax tables;
// the formulas were checked against torch.autograd to 1e-15 in fp64 (tests/test_14_meta_grad_gpu.py
// checks this kernel against the autograd oracle).
//
// Every row scatter is a segmented sum over the row-sorted token list: deterministic, no atomics.
#include "lpg_common.cuh"
#include "segreduce.cuh"
#include "../../include/toued.h"

// ------------------------------------------------------------------------------------------------
// meta loss on the eval rollout + lam_K
__global__ void __launch_bounds__(256)
meta_loss_kernel(const int32_t* __restrict__ obs, const uint8_t* __restrict__ action,
                 const float* __restrict__ reward, const uint8_t* __restrict__ done,
                 const uint16_t* __restrict__ sorted_tok, const float* __restrict__ value,
                 const float* __restrict__ actor, float* __restrict__ lam, float* __restrict__ mu,
                 float* __restrict__ scal, int W, int L, int D, int vstride, float gamma, float lmbda,
                 float gscale, int quirk) {
    extern __shared__ __align__(16) float sm[];
    float* adv = sm;                       // [L][W]
    float* rec = adv + L * W;              // [T][6]: 5 dlogits + tf
    float* abar = rec + L * W * 6;         // [W]
    float* lbar = abar + W;                // [W]
    float* runv = lbar + W;                // [T][5]
    float* scan = runv + L * W * 5;        // [2][256][5]
    void* idxmem = scan + 2 * 256 * 5;
    __shared__ float red[32];
    __shared__ int iscan[512];
    __shared__ unsigned char sflags[512];
    const int n = blockIdx.x, tid = threadIdx.x, T = W * L;
    const int32_t* ob = obs + (size_t)n * (L + 1) * W;
    const uint8_t* act = action + (size_t)n * T;
    const float* rw = reward + (size_t)n * T;
    const uint8_t* dn = done + (size_t)n * T;
    const uint16_t* st = sorted_tok + (size_t)n * T;
    const float* vt = value + (size_t)n * D * vstride;
    const float* at = actor + (size_t)n * D * 8;
    float* lm = lam + (size_t)n * D * 8;
    float* mm = mu + (size_t)n * D * 8;
    const float invT = 1.0f / (float)T;
    const SegIndex si = seg_index_build(idxmem, iscan, st, ob, T);

    for (int i = tid; i < D * 2; i += 256) {
        reinterpret_cast<float4*>(lm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(mm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- GAE per worker (util/metrics.py:17-38); value = W_v[row] + 0.001 t W_v[D-1] ----
    float s1 = 0.f, s2 = 0.f;
    const float vlast = vt[(size_t)(D - 1) * vstride];
    for (int w = tid; w < W; w += 256) {
        float g = 0.f;
        int32_t o1 = ob[L * W + w];
        float v1 = vt[(size_t)ob_idx(o1) * vstride] + 0.001f * (float)ob_time(o1) * vlast;
        for (int t = L - 1; t >= 0; --t) {
            const int32_t o0 = ob[t * W + w];
            const float v0 = vt[(size_t)ob_idx(o0) * vstride] + 0.001f * (float)ob_time(o0) * vlast;
            const float nd = dn[t * W + w] ? 0.0f : 1.0f;
            const float delta = rw[t * W + w] + gamma * v1 * nd - v0;
            g = delta + gamma * lmbda * nd * g;
            adv[t * W + w] = g;
            s1 += g; s2 = fmaf(g, g, s2);          // target - value == adv  ->  value_loss = mean(adv^2)
            v1 = v0;
        }
    }
    s1 = block_sum(s1, red); s2 = block_sum(s2, red);
    const float mean = s1 * invT;
    float var = 0.f;
    for (int i = tid; i < T; i += 256) { const float d = adv[i] - mean; var = fmaf(d, d, var); }
    var = block_sum(var, red) * invT;
    const float inv_std = 1.0f / (sqrtf(var) + 1e-8f);                   // meta/train.py:85
    __syncthreads();
    for (int i = tid; i < T; i += 256) adv[i] = (adv[i] - mean) * inv_std;
    __syncthreads();
    // ---- per-token log-prob of the sampled action under theta_K, dlogits ----
    for (int tok = tid; tok < T; tok += 256) {
        const int32_t o = ob[tok];
        float z[5], p[5];
        tab_logits8<5>(at, D, o, z);
        softmax_c<5>(z, p);
        const int a = act[tok];
        float pa = p[0];
#pragma unroll
        for (int j = 1; j < 5; ++j) pa = (a == j) ? p[j] : pa;
        float* r = rec + tok * 6;
        r[5] = logf(pa + 1e-8f);                                         // temporarily: log-prob
        const float q = pa / (pa + 1e-8f);
#pragma unroll
        for (int j = 0; j < 5; ++j) r[j] = q * ((a == j ? 1.0f : 0.0f) - p[j]);
    }
    __syncthreads();
    float loss = 0.f;
    for (int w = tid; w < W; w += 256) {
        float sa = 0.f, sl = 0.f, se = 0.f;
        for (int t = 0; t < L; ++t) { sa += adv[t * W + w]; sl += rec[(t * W + w) * 6 + 5]; se = fmaf(adv[t * W + w], rec[(t * W + w) * 6 + 5], se); }
        abar[w] = sa / (float)L; lbar[w] = sl / (float)L;
        // Q16: mean over the [L, L] outer product = mean_t(adv) * mean_t(logp) ; else element-wise mean
        loss -= quirk ? abar[w] * lbar[w] : se / (float)L;
    }
    loss = block_sum(loss, red) / (float)W;
    __syncthreads();
    // coefficient of grad log(pi+1e-8)[tok] in d(lpg_loss)/d theta_K, times 1/N_global
    float last[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int tok = tid; tok < T; tok += 256) {
        const int w = tok % W;
        const float coef = -gscale * invT * (quirk ? abar[w] : adv[tok]);
        float* r = rec + tok * 6;
        const float tf = 0.001f * (float)ob_time(ob[tok]);
#pragma unroll
        for (int j = 0; j < 5; ++j) { r[j] *= coef; last[j] = fmaf(tf, r[j], last[j]); }
        r[5] = tf;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) last[j] = block_sum(last[j], red);
    __syncthreads();
    seg_reduce<5, 6>(rec, si, T, runv, scan, sflags);
    for (int r = tid; r < si.nruns; r += 256) {
        const int row = si.run_row[r];
#pragma unroll
        for (int j = 0; j < 5; ++j) lm[(size_t)row * 8 + j] = runv[r * 5 + j];
    }
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) lm[(size_t)(D - 1) * 8 + j] = last[j];
        scal[n * 2 + 0] = loss;               // lpg_loss
        scal[n * 2 + 1] = s2 * invT;          // value_loss
    }
}

extern "C" int toued_meta_loss(const int32_t* obs, const uint8_t* action, const float* reward,
                               const uint8_t* done, const uint16_t* sorted_tok, const float* value_table,
                               const float* actor, float* lam, float* mu, float* scalars, int n_agents,
                               int n_workers, int rollout_len, int obs_dim, int value_stride, float gamma,
                               float gae_lambda, float grad_scale, int outer_product_quirk, void* stream) {
    const int T = n_workers * rollout_len;
    TOUED_CHECK(n_agents > 0 && T > 0, "toued_meta_loss: empty problem");
    const size_t smem = sizeof(float) * ((size_t)T * 7 + 2 * n_workers + (size_t)T * 5 + 2 * 256 * 5) + seg_index_bytes(T);
    TOUED_CHECK(smem <= 200 * 1024, "toued_meta_loss: W*L=%d too large for shared memory", T);
    TOUED_CUDA(cudaFuncSetAttribute(meta_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    meta_loss_kernel<<<n_agents, 256, smem, (cudaStream_t)stream>>>(
        obs, action, reward, done, sorted_tok, value_table, actor, lam, mu, scalars, n_workers, rollout_len,
        obs_dim, value_stride, gamma, gae_lambda, grad_scale, outer_product_quirk);
    TOUED_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// adjoint of agent update k
struct TokFwd { float p[5], y[8], pa, q, tf; int a; };

__device__ __forceinline__ TokFwd tok_forward(const float* at, const float* ct, int D, int32_t o, int a) {
    TokFwd f;
    float z[5], zy[8];
    tab_logits8<5>(at, D, o, z);
    softmax_c<5>(z, f.p);
    tab_logits8<8>(ct, D, o, zy);
    softmax_c<8>(zy, f.y);
    f.a = a;
    f.pa = f.p[0];
#pragma unroll
    for (int j = 1; j < 5; ++j) f.pa = (a == j) ? f.p[j] : f.pa;
    f.q = f.pa / (f.pa + 1e-8f);
    f.tf = 0.001f * (float)ob_time(o);
    return f;
}

// forward-gradient dlogits of token tok at (theta_k, phi_k): c[0..5) actor, c[5..13) critic
__device__ __forceinline__ void tok_grad(const TokFwd& f, float ph, const float* yh, float invT, float alpha, float* c) {
    const float ca = invT * ph * f.q;
#pragma unroll
    for (int j = 0; j < 5; ++j) c[j] = ca * ((f.a == j ? 1.0f : 0.0f) - f.p[j]);
    float mm[8], b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mm[i] = logf(f.y[i] + 1e-8f) - logf(yh[i] + 1e-8f) + f.y[i] / (f.y[i] + 1e-8f);
        b = fmaf(f.y[i], mm[i], b);
    }
    const float cc = alpha * invT;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[5 + i] = cc * f.y[i] * (mm[i] - b);
}

__device__ __forceinline__ void load8(const float* p, float* v) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(256, 2)
agent_backward_kernel(const int32_t* __restrict__ obs, const uint8_t* __restrict__ action,
                      const uint16_t* __restrict__ sorted_tok, const float* __restrict__ pi_hat,
                      const float* __restrict__ y_hat, const float* __restrict__ actor_k,
                      const float* __restrict__ critic_k, const float* __restrict__ actor_k1,
                      const float* __restrict__ critic_k1, const float* __restrict__ upd_scal,
                      float* lam, float* mu, float* __restrict__ d_pi_hat, float* __restrict__ d_y_hat,
                      float* run_scratch, int n_agents, int W, int L, int D, float lr_a, float lr_c, float max_norm,
                      float alpha, float b_pent, float b_yent, float b_pl2, float b_yl2, float gscale) {
    extern __shared__ __align__(16) float rec[];       // [T][13] records | scan | index   (106 KB at T = 1280: two CTAs per SM)
    __shared__ float red[32];
    __shared__ float s_wlast[13];
    __shared__ int iscan[512];
    __shared__ unsigned char sflags[512];
    const int n = blockIdx.x, tid = threadIdx.x, T = W * L, R = n_agents * W;
    const int32_t* ob = obs + (size_t)n * (L + 1) * W;
    const uint8_t* act = action + (size_t)n * T;
    const uint16_t* st = sorted_tok + (size_t)n * T;
    const float* a0 = actor_k + (size_t)n * D * 8;
    const float* c0 = critic_k + (size_t)n * D * 8;
    const float* a1 = actor_k1 + (size_t)n * D * 8;
    const float* c1 = critic_k1 + (size_t)n * D * 8;
    float* lm = lam + (size_t)n * D * 8;
    float* mm = mu + (size_t)n * D * 8;
    const float invT = 1.0f / (float)T;
    const float gna = upd_scal[n * 8 + 0], gnc = upd_scal[n * 8 + 1];
    const bool keep = upd_scal[n * 8 + 2] != 0.0f;
    // run vectors [min(T, D)][13] (a run is a distinct table row: at most D of them) live in an L2-resident global
    // scratch: written once per run, read once per run / token
    float* runv = run_scratch + (size_t)n * min(T, D) * 13;
    float* scan = rec + (size_t)T * 13;
    const SegIndex si = seg_index_build(scan + 2 * 256 * 13, iscan, st, ob, T);

    // ---- A. entropy regularisers evaluated at the UPDATED tables (lpg_agent.py:119-120) -------
    // L has  -b_pent/K * H(pi_{theta_{k+1}})  and  -b_yent/K * H(y_{phi_{k+1}})   (1/K folded into b_*)
    float last[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) last[j] = 0.f;
    for (int tok = tid; tok < T; tok += 256) {
        const TokFwd f = tok_forward(a1, c1, D, ob[tok], act[tok]);
        float* r = rec + tok * 13;
        {   // dH/dp_j = -(log(p_j + e) + 1);  dz = p * (dHp - sum p dHp)
            float dh[5], s = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) { dh[j] = -(logf(f.p[j] + 1e-8f) + 1.0f); s = fmaf(f.p[j], dh[j], s); }
            const float k = -b_pent * gscale * invT;
#pragma unroll
            for (int j = 0; j < 5; ++j) r[j] = k * f.p[j] * (dh[j] - s);
        }
        {
            float dh[8], s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { dh[j] = -(logf(f.y[j] + 1e-8f) + 1.0f); s = fmaf(f.y[j], dh[j], s); }
            const float k = -b_yent * gscale * invT;
#pragma unroll
            for (int j = 0; j < 8; ++j) r[5 + j] = k * f.y[j] * (dh[j] - s);
        }
#pragma unroll
        for (int j = 0; j < 13; ++j) last[j] = fmaf(f.tf, r[j], last[j]);
    }
#pragma unroll
    for (int j = 0; j < 13; ++j) last[j] = block_sum(last[j], red);
    __syncthreads();
    seg_reduce<13, 13>(rec, si, T, runv, scan, sflags);
    for (int r = tid; r < si.nruns; r += 256) {
        const int row = si.run_row[r];
        const float* g = runv + r * 13;
#pragma unroll
        for (int j = 0; j < 5; ++j) lm[(size_t)row * 8 + j] += g[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) mm[(size_t)row * 8 + j] += g[5 + j];
    }
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) lm[(size_t)(D - 1) * 8 + j] += last[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) mm[(size_t)(D - 1) * 8 + j] += last[5 + j];
    }
    __syncthreads();

    const float dir_pi = gscale * b_pl2 * 2.0f * invT;      // d/d pi_hat of (b_pl2/K) mean(pi_hat^2)
    const float dir_y = gscale * b_yl2 * 2.0f * invT;       // d/d y_hat  of (b_yl2/K) mean(sum y_hat^2)
    if (!keep) {                                             // masked update: theta_{k+1} = theta_k
        for (int tok = tid; tok < T; tok += 256) {
            const int t = tok / W, w = tok - t * W;
            const size_t li = (size_t)t * R + (size_t)n * W + w;
            d_pi_hat[li] = dir_pi * pi_hat[li];
            float yh[8]; load8(y_hat + li * 8, yh);
            float4* o = reinterpret_cast<float4*>(d_y_hat + li * 8);
            o[0] = make_float4(dir_y * yh[0], dir_y * yh[1], dir_y * yh[2], dir_y * yh[3]);
            o[1] = make_float4(dir_y * yh[4], dir_y * yh[5], dir_y * yh[6], dir_y * yh[7]);
        }
        return;
    }

    // ---- B1. forward-gradient dlogits at (theta_k, phi_k); time-row gradient -------------------
#pragma unroll
    for (int j = 0; j < 13; ++j) last[j] = 0.f;
    for (int tok = tid; tok < T; tok += 256) {
        const int t = tok / W, w = tok - t * W;
        const size_t li = (size_t)t * R + (size_t)n * W + w;
        const TokFwd f = tok_forward(a0, c0, D, ob[tok], act[tok]);
        float yh[8]; load8(y_hat + li * 8, yh);
        float* r = rec + tok * 13;
        tok_grad(f, pi_hat[li], yh, invT, alpha, r);
#pragma unroll
        for (int j = 0; j < 13; ++j) last[j] = fmaf(f.tf, r[j], last[j]);
    }
#pragma unroll
    for (int j = 0; j < 13; ++j) last[j] = block_sum(last[j], red);     // g_k[D-1]
    __syncthreads();
    // ---- B2. <g_k, lam>, <g_k^c, mu> ------------------------------------------------------------
    seg_reduce<13, 13>(rec, si, T, runv, scan, sflags);           // g_k rows
    float dot_a = 0.f, dot_c = 0.f;
    for (int r = tid; r < si.nruns; r += 256) {
        const int row = si.run_row[r];
        const float* g = runv + r * 13;
        float l8[8], m8[8];
        load8(lm + (size_t)row * 8, l8); load8(mm + (size_t)row * 8, m8);
#pragma unroll
        for (int j = 0; j < 5; ++j) dot_a = fmaf(g[j], l8[j], dot_a);
#pragma unroll
        for (int j = 0; j < 8; ++j) dot_c = fmaf(g[5 + j], m8[j], dot_c);
    }
    dot_a = block_sum(dot_a, red); dot_c = block_sum(dot_c, red);
    float ll[8], ml[8];
    load8(lm + (size_t)(D - 1) * 8, ll); load8(mm + (size_t)(D - 1) * 8, ml);
#pragma unroll
    for (int j = 0; j < 5; ++j) dot_a = fmaf(last[j], ll[j], dot_a);
#pragma unroll
    for (int j = 0; j < 8; ++j) dot_c = fmaf(last[5 + j], ml[j], dot_c);
    // w = -lr * J_C^T lam :  unclipped -> -lr*lam ; clipped -> -lr*(c/|g|) * (lam - g <g,lam>/|g|^2)
    const bool clip_a = !(gna < max_norm), clip_c = !(gnc < max_norm);
    const float ka = -lr_a * (clip_a ? max_norm / gna : 1.0f), pa_ = clip_a ? dot_a / (gna * gna) : 0.0f;
    const float kc = -lr_c * (clip_c ? max_norm / gnc : 1.0f), pc_ = clip_c ? dot_c / (gnc * gnc) : 0.0f;
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) s_wlast[j] = ka * (ll[j] - pa_ * last[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s_wlast[5 + j] = kc * (ml[j] - pc_ * last[5 + j]);
    }
    __syncthreads();
    float wl[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) wl[j] = s_wlast[j];

    // ---- B3a. per row: w_row = k * (lam_row - proj * g_row), kept as the run's vector -----------------
    for (int r = tid; r < si.nruns; r += 256) {
        const int row = si.run_row[r];
        float* g = runv + r * 13;
        float l8[8], m8[8];
        load8(lm + (size_t)row * 8, l8); load8(mm + (size_t)row * 8, m8);
#pragma unroll
        for (int j = 0; j < 5; ++j) g[j] = ka * (l8[j] - pa_ * g[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[5 + j] = kc * (m8[j] - pc_ * g[5 + j]);
    }
    __syncthreads();
    // ---- B3b. token-parallel: cotangents of pi_hat / y_hat and the Hessian-vector contributions ----
    float hl[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) hl[j] = 0.f;
    for (int pos = tid; pos < T; pos += 256) {
        const int tok = si.tok[pos];
        const int t = tok / W, w = tok - t * W;
        const size_t li = (size_t)t * R + (size_t)n * W + w;
        const TokFwd f = tok_forward(a0, c0, D, ob[tok], act[tok]);
        const float ph = pi_hat[li];
        float yh[8]; load8(y_hat + li * 8, yh);
        float* r = rec + tok * 13;
        const float* wr = runv + (size_t)si.run[pos] * 13;
        float v[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) v[j] = fmaf(f.tf, wl[j], wr[j]);
        // actor:  S = q * (v_a - p.v)
        float va = v[0], pv = 0.f;
#pragma unroll
        for (int j = 1; j < 5; ++j) va = (f.a == j) ? v[j] : va;
#pragma unroll
        for (int j = 0; j < 5; ++j) pv = fmaf(f.p[j], v[j], pv);
        const float s = va - pv;
        d_pi_hat[li] = invT * f.q * s + dir_pi * ph;
        const float pe = f.pa + 1e-8f;
        const float c1_ = ph * invT * s * 1e-8f * f.pa / (pe * pe), c2_ = ph * invT * f.q;
        float h[13];
#pragma unroll
        for (int j = 0; j < 5; ++j)
            h[j] = c1_ * ((f.a == j ? 1.0f : 0.0f) - f.p[j]) - c2_ * f.p[j] * (v[j] - pv);
        // critic:  S = sum_i v_i y_i m_i - (v.y)(y.m)
        float m[8], mp[8], av = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float ye = f.y[j] + 1e-8f;
            m[j] = logf(ye) - logf(yh[j] + 1e-8f) + f.y[j] / ye;
            mp[j] = (f.y[j] + 2e-8f) / (ye * ye);
            av = fmaf(v[5 + j], f.y[j], av);
            b = fmaf(f.y[j], m[j], b);
        }
        const float ca = alpha * invT;
        float dS[8], sy = 0.f, dyo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float vc = v[5 + j];
            dyo[j] = ca * f.y[j] * (av - vc) / (yh[j] + 1e-8f) + dir_y * yh[j];
            dS[j] = vc * m[j] + vc * f.y[j] * mp[j] - vc * b - av * (m[j] + f.y[j] * mp[j]);
            sy = fmaf(f.y[j], dS[j], sy);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) h[5 + j] = ca * f.y[j] * (dS[j] - sy);
        float4* o = reinterpret_cast<float4*>(d_y_hat + li * 8);
        o[0] = make_float4(dyo[0], dyo[1], dyo[2], dyo[3]);
        o[1] = make_float4(dyo[4], dyo[5], dyo[6], dyo[7]);
#pragma unroll
        for (int j = 0; j < 13; ++j) { r[j] = h[j]; hl[j] = fmaf(f.tf, h[j], hl[j]); }
    }
#pragma unroll
    for (int j = 0; j < 13; ++j) hl[j] = block_sum(hl[j], red);
    __syncthreads();
    // ---- B3c. segmented sums of the HVP contributions into lam_k / mu_k --------------------------
    seg_reduce<13, 13>(rec, si, T, runv, scan, sflags);
    for (int r = tid; r < si.nruns; r += 256) {
        const int row = si.run_row[r];
        const float* g = runv + r * 13;
#pragma unroll
        for (int j = 0; j < 5; ++j) lm[(size_t)row * 8 + j] += g[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) mm[(size_t)row * 8 + j] += g[5 + j];
    }
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) lm[(size_t)(D - 1) * 8 + j] = ll[j] + hl[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) mm[(size_t)(D - 1) * 8 + j] = ml[j] + hl[5 + j];
    }
}

extern "C" int toued_agent_backward(const int32_t* obs, const uint8_t* action, const uint16_t* sorted_tok,
                                    const float* pi_hat, const float* y_hat, const float* actor_k,
                                    const float* critic_k, const float* actor_k1, const float* critic_k1,
                                    const float* update_scalars, float* lam, float* mu, float* d_pi_hat,
                                    float* d_y_hat, int n_agents, int n_workers, int rollout_len, int obs_dim,
                                    float lr_actor, float lr_critic, float max_grad_norm,
                                    float agent_target_coeff, float policy_entropy_coeff,
                                    float target_entropy_coeff, float policy_l2_coeff, float target_l2_coeff,
                                    float grad_scale, float* run_scratch, void* stream) {
    const int T = n_workers * rollout_len;
    TOUED_CHECK(n_agents > 0 && T > 0, "toued_agent_backward: empty problem");
    TOUED_CHECK(run_scratch != nullptr, "toued_agent_backward: run_scratch (toued_agent_scratch_floats) is required");
    const size_t smem = sizeof(float) * ((size_t)T * 13 + 2 * 256 * 13) + seg_index_bytes(T);
    TOUED_CHECK(smem <= 200 * 1024, "toued_agent_backward: W*L=%d too large for shared memory", T);
    TOUED_CUDA(cudaFuncSetAttribute(agent_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TOUED_CUDA(cudaFuncSetAttribute(agent_backward_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cudaStream_t st = (cudaStream_t)stream;
    agent_backward_kernel<<<n_agents, 256, smem, st>>>(
        obs, action, sorted_tok, pi_hat, y_hat, actor_k, critic_k, actor_k1, critic_k1, update_scalars, lam, mu,
        d_pi_hat, d_y_hat, run_scratch, n_agents, n_workers, rollout_len, obs_dim, lr_actor, lr_critic, max_grad_norm,
        agent_target_coeff, policy_entropy_coeff, target_entropy_coeff, policy_l2_coeff, target_l2_coeff, grad_scale);
    TOUED_LAUNCH_CHECK();
    return 0;
}
