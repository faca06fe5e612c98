// Tensor-core GRU backward (BPTT) for the LPG network — the reverse of gru_forward_tc.cu, i.e. the part
// of jax.grad(_train_agent) (meta/train.py:121-128) that flows through models/lpg.py:11-30,77-84.
//
// A CTA owns 128 sequences and walks t = 0 .. L-1 (the forward scan is reversed).  Per step
//     dh_t   = relu'(h_t) * (d pi_hat * w_pi + dl . W_y)  +  (1 - done_{t-1}) * carry
//     dG_g   = dh_t * f_g            (f_r, f_z, f_hn, f_an saved by the forward epilogue)
//     carry' = dGh @ Wh^T + z * dh_t
// The carry stays in fp32 in TMEM (two 256-column sets, ping-pong): dGh @ Wh^T is accumulated by
// tcgen05.mma from fp16 operands (dG written by the epilogue warps into SW128 smem chunks, Wh chunks
// streamed by a TMA-producer warp), and the element-wise z*dh_t term is injected into the same
// accumulator by a tiny identity MMA on an fp16 hi/lo pair, so the recurrent chain keeps ~22 mantissa
// bits without ever leaving tensor memory.
// Scaled arithmetic (tc.cuh, "scaled fp16 operands"): the kernel multiplies the incoming cotangents by the power of two
// S derived from `cotmax` and everything downstream -- carry, dG, dl, dx -- is in units of S; the consumers
// (toued_lpg_wgrad_tc, toued_lpg_wgrad_embed) take S back out.  Round 1 used bf16 operands (8 significant bits; the
// rounding of Wh is systematic over all tokens) and left 0.6-1.6e-2 of error on the GRU weight blocks of the meta-gradient.
// Outputs for the weight-gradient kernels: S * dG (dar, daz, dhn, dan) as an fp16 token-tile image,
// the head-logit cotangents S * dl and S * (d pyt, d pyt1).
#include "tc.cuh"
#include "lpg_common.cuh"
#include "../../include/toued.h"

constexpr int BT_M = 128;
// Epilogue warps: warp w owns TMEM lanes [32 (w & 3), +32) and the unit slice (w >> 2) of every 64-unit block.
// BT_EW = 8 (two 32-unit halves, 168 registers) is the production geometry.  BT_EW = 16 (four 16-unit quarters, one K-step
// of the A stage each; library variant "bwd16") was measured in round 2: correct, but 5.96 instead of 4.24 ms per meta-step
// -- at 608 threads the kernel is capped at 96 registers and spills 284 B per thread in the chunk loop, and four warps per
// TMEM quadrant queue on the same tcgen05.ld port.  More warps do not buy latency hiding here; fewer instructions would.
#ifndef BT_EW
#define BT_EW 8
#endif
constexpr int BT_NQ = BT_EW / 4;                 // unit slices per 64-unit block (2 halves or 4 quarters)
constexpr int BT_NC = 8 / BT_NQ;                 // 8-unit chunks per thread and unit block (4 or 2)
constexpr int BT_THREADS = (BT_EW + 3) * 32;     // epilogue warps + TMA-load warp + MMA warp + TMA-store warp
static_assert(BT_EW == 8 || BT_EW == 16, "8 or 16 epilogue warps");
constexpr int BT_ACHUNK = BT_M * 128;            // 16 KB: [128 rows][64 bf16]
constexpr int BT_ASTAGE = 6 * BT_ACHUNK;         // dG_r, dG_z, dG_hn, cz_hi, cz_lo, dG_an (store only)
constexpr int BT_BCHUNK = LPG_H * 128;           // 32 KB: [256 units][64 c]
constexpr int BT_NSB = 3;
constexpr int BT_ICHUNK = 64 * 128;              // 8 KB identity

// Wh[j][c] -> 12 fp16 K-major SW128 chunk images [cc = c / 64][row j][k = c % 64]
__global__ void pack_wh_bwd_kernel(const float* __restrict__ Wh, __half* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LPG_H * LPG_G) return;
    const int j = i / LPG_G, c = i % LPG_G;
    char* base = reinterpret_cast<char*>(img) + (size_t)(c >> 6) * BT_BCHUNK;
    *reinterpret_cast<__half*>(base + sw128_offset(LPG_H, j, c & 63)) = __float2half_rn(Wh[i]);
}

extern "C" int toued_pack_wh_backward(const float* lpg_params, void* whb_img, void* stream) {
    pack_wh_bwd_kernel<<<(LPG_H * LPG_G + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        lpg_params + lpg_offsets(5).Wh, (__half*)whb_img);
    TOUED_LAUNCH_CHECK();
    return 0;
}

__device__ __forceinline__ void unpack8h(const uint4& r, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}
__global__ void __launch_bounds__(BT_THREADS, 1)
gru_backward_tc_kernel(const uint8_t* __restrict__ done, const float* __restrict__ lpg, int X,
                       const unsigned char* __restrict__ whb_img, const __half* __restrict__ h16,
                       const __half* __restrict__ fac, const float* __restrict__ y_hat,
                       const float* __restrict__ d_pi_hat, const float* __restrict__ d_y_hat,
                       unsigned char* __restrict__ dgimg, float* __restrict__ dl_out, float* __restrict__ dx,
                       const uint32_t* __restrict__ cotmax, int R, int L, int W) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // (offset arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the shared
    //  address space and emits LDS / STS instead of generic LD / ST)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = smem;                                   // 96 KB
    unsigned char* sB = sA + BT_ASTAGE;                         // 3 x 32 KB
    unsigned char* sI = sB + BT_NSB * BT_BCHUNK;                // 8 KB identity
    float* swp = reinterpret_cast<float*>(sI + BT_ICHUNK);      // [256]
    float* sWy = swp + LPG_H;                                   // [256][8]
    float* sWi = sWy + LPG_H * LPG_Y;                           // [256 units][8]: Wi rows 3, 4 x gates (r, z, n), 2 pad
    float* sdx = sWi + LPG_H * 8;                               // [BT_NQ - 1][128][2]
    unsigned char* ssign = reinterpret_cast<unsigned char*>(sdx + BT_M * 2 * (BT_NQ - 1));   // [4 BT_NC chunks][epilogue threads]: relu'(h_t) bits
    __shared__ __align__(8) uint64_t b_full[BT_NSB], b_empty[BT_NSB], k_full[4], a_empty, q_full;
    __shared__ uint32_t tmem_base_s;

    const LpgOffsets o = lpg_offsets(X);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * BT_M;

    if (tid == 0) {
        for (int s = 0; s < BT_NSB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int k = 0; k < 4; ++k) mbar_init(&k_full[k], 4);
        mbar_init(&a_empty, 2); mbar_init(&q_full, 1);
        mbar_fence_init();
    }
    if (warp == BT_EW + 1) tmem_alloc(&tmem_base_s, 512);
    for (int i = tid; i < LPG_H; i += BT_THREADS) swp[i] = lpg[o.w_pi + i];
    for (int i = tid; i < LPG_H * LPG_Y; i += BT_THREADS) sWy[i] = lpg[o.W_y + i];
    for (int i = tid; i < LPG_H * 8; i += BT_THREADS) {
        const int u = i >> 3, q = i & 7;               // q = 3 * (input row - 3) + gate
        sWi[i] = q < 6 ? lpg[o.Wi + (3 + q / 3) * LPG_G + (q % 3) * LPG_H + u] : 0.0f;
    }
    for (int i = tid; i < 64 * 64; i += BT_THREADS) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<__half*>(sI + sw128_offset(64, n, k)) = __float2half_rn(n == k ? 1.0f : 0.0f);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == BT_EW) {
        // ===================== TMA producer: Wh chunks in (unit block, gate) order ====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t + 1 < L; ++t)
                for (int ub = 0; ub < 4; ++ub)
                    for (int g = 0; g < 3; ++g, ++it) {
                        const int s = it % BT_NSB;
                        mbar_wait(&b_empty[s], ((it / BT_NSB) & 1) ^ 1);
                        mbar_expect_tx(&b_full[s], BT_BCHUNK);
                        bulk_g2s(sB + s * BT_BCHUNK, whb_img + (size_t)(g * 4 + ub) * BT_BCHUNK, BT_BCHUNK, &b_full[s]);
                    }
        }
    } else if (warp == BT_EW + 1) {
        // ===================== MMA issuer =============================================================
        // The whole warp walks the loop (all lanes wait on the barriers); one elected lane issues (tc.cuh::elect_one).
        {
            constexpr uint32_t idesc256 = tc_idesc(BT_M, 256, 0), idesc64 = tc_idesc(BT_M, 64, 0);     // fp16 operands
            const uint32_t a_addr = smem_u32(sA), i_addr = smem_u32(sI);
            const uint64_t ad0 = tc_smem_desc(a_addr), idd = tc_smem_desc(i_addr);
            uint32_t it = 0;
            uint32_t ait = 0;
            for (int t = 0; t < L; ++t) {
                const uint32_t q_addr = tmem_base + (t & 1) * 256;
                const bool mma = t + 1 < L;                          // the last step only stores its dG tiles
                for (int ub = 0; ub < 4; ++ub, ++ait) {
                    // The A stage is consumed K-step by K-step (16 units = what four epilogue warps finish every two
                    // chunks): the MMAs of a unit block are spread over the time the epilogue warps need to produce
                    // it, so that only the last quarter is still outstanding when they want the stage back.
                    uint32_t sg[3];
                    if (mma) {
#pragma unroll
                        for (int g = 0; g < 3; ++g, ++it) {
                            sg[g] = it % BT_NSB;
                            mbar_wait(&b_full[sg[g]], (it / BT_NSB) & 1);
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        // K-steps in the order the epilogue completes them: with two unit halves 0, 2, 1, 3 (the halves
                        // progress together), with four quarters 0, 1, 2, 3 (all four finish at the same time)
                        const int ks = BT_NQ == 2 ? (((kk & 1) << 1) | (kk >> 1)) : kk;
                        mbar_wait(&k_full[ks], ait & 1);
                        tc_fence_after();
                        if (mma && elect_one()) {
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                                tc_mma(q_addr, ad0 + (uint64_t)((g * BT_ACHUNK) >> 4) + 2 * ks,
                                       tc_smem_desc(smem_u32(sB + sg[g] * BT_BCHUNK)) + 2 * ks, idesc256, (ub | kk | g) != 0);
                            // z * dh_t (bf16 hi + lo) through the identity: accumulates onto columns [64 ub, 64 ub + 64)
                            tc_mma(q_addr + ub * 64, ad0 + (uint64_t)((3 * BT_ACHUNK) >> 4) + 2 * ks, idd + 2 * ks, idesc64, 1u);
                            tc_mma(q_addr + ub * 64, ad0 + (uint64_t)((4 * BT_ACHUNK) >> 4) + 2 * ks, idd + 2 * ks, idesc64, 1u);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        if (mma) {
#pragma unroll
                            for (int g = 0; g < 3; ++g) tc_commit(&b_empty[sg[g]]);
                        }
                        tc_commit(&a_empty);                         // arrival 1 of 2: the MMAs have read the stage
                        if (mma && ub == 3) tc_commit(&q_full);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == BT_EW + 2) {
        // ===================== TMA store warp: dG tiles of every unit block -> token-tile image ==========
        // The SW128 chunks in smem ARE the image's 8 KB sub-tiles of the weight-gradient GEMM, so they leave as
        // full-line bulk stores.  A warp of its own: waiting for the stores to finish reading the stage must not
        // hold up the MMA issue of the next unit block.
        if (lane == 0) {
            const uint32_t a_addr = smem_u32(sA);
            const size_t Rp = ((size_t)R + 63) & ~(size_t)63;
            const int nsub = ((size_t)row0 + 64 < Rp) ? 2 : 1;      // 64-token sub-tiles of this CTA inside the image
            uint32_t ait = 0;
            for (int t = 0; t < L; ++t) {
                for (int ub = 0; ub < 4; ++ub, ++ait) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mbar_wait(&k_full[ks], ait & 1);
                    const size_t itok0 = (size_t)t * Rp + row0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t src = a_addr + (g == 3 ? 5 : g) * BT_ACHUNK;
                        for (int sub = 0; sub < nsub; ++sub)
                            bulk_s2g(dgimg + tile_img_offset(itok0 + sub * 64, 16, (g * 4 + ub) * 64), src + sub * 8192, 8192);
                    }
                    bulk_commit();
                    bulk_wait_read();
                    mbar_arrive(&a_empty);                           // arrival 2 of 2: the stores have read the stage
                }
            }
            bulk_wait_all();
        }
    } else {
        // ===================== epilogue / producer-of-A warps 0 .. BT_EW-1 ==============================
        const int q = warp & 3, hf = warp >> 2;
        const int rl = q * 32 + lane;
        const int row = row0 + rl;
        const bool rv = row < R;
        const int rsafe = rv ? row : 0;
        const int n_ag = rsafe / W, w_ag = rsafe % W;
        const size_t R32 = ((size_t)R + 31) >> 5;
        const size_t gs = (size_t)L * R32 * 32 * LPG_H;
        uint32_t ait = 0;
        const uint32_t sA_u32 = smem_u32(sA);
        const float S = cotmax ? cot_scale_from_max(*cotmax) : 1.0f;      // power of two: scaling is exact
        // ---- software pipeline: the factor loads of the next 8-unit chunk and the head cotangents of the
        //      next timestep are issued one iteration ahead (across unit-block / timestep boundaries) ----
        struct FacLoads { uint4 r, z, n, hn, hx; };              // gates of step t, h of step t+1 (the carry h')
        const size_t tstride = R32 * 32 * LPG_H;                 // RB32 elements per timestep
        auto issue_fac = [&](int t_, int ub_, int c8_) {
            FacLoads l;
            const size_t base = rb32_index((size_t)t_, R32, rsafe, ub_ * 64 + hf * (64 / BT_NQ) + c8_ * 8);
            l.r = *reinterpret_cast<const uint4*>(fac + base);
            l.z = *reinterpret_cast<const uint4*>(fac + gs + base);
            l.n = *reinterpret_cast<const uint4*>(fac + 2 * gs + base);
            l.hn = *reinterpret_cast<const uint4*>(fac + 3 * gs + base);
            l.hx = t_ + 1 < L ? *reinterpret_cast<const uint4*>(h16 + base + tstride) : make_uint4(0u, 0u, 0u, 0u);
            return l;
        };
        struct RowLoads { float4 y0, y1, d0, d1; float dpi; uint8_t dn, dn_hp; };
        auto issue_row = [&](int t_) {
            RowLoads r;
            const size_t tok_ = (size_t)t_ * R + rsafe;
            const float4* q0 = reinterpret_cast<const float4*>(y_hat + tok_ * 8);
            const float4* q1 = reinterpret_cast<const float4*>(d_y_hat + tok_ * 8);
            r.y0 = q0[0]; r.y1 = q0[1]; r.d0 = q1[0]; r.d1 = q1[1];
            r.dpi = d_pi_hat[tok_];
            r.dn = t_ > 0 ? done[((size_t)n_ag * L + (t_ - 1)) * W + w_ag] : (uint8_t)1;
            r.dn_hp = t_ + 1 < L ? done[((size_t)n_ag * L + t_) * W + w_ag] : (uint8_t)1;   // the cell at t consumed (1 - done_t) h_{t+1}
            return r;
        };
        // relu'(h_0) bits of this thread's 4 BT_NC chunks (later steps get theirs from the h' load one step earlier)
        const int et = tid;                                      // index among the epilogue threads
        constexpr int ET = BT_EW * 32;
        for (int i = 0; i < 4 * BT_NC; ++i) {
            const uint4 raw = *reinterpret_cast<const uint4*>(h16 + rb32_index(0, R32, rsafe, (i / BT_NC) * 64 + hf * (64 / BT_NQ) + (i % BT_NC) * 8));
            float v[8];
            unpack8h(raw, v);
            unsigned b = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) b |= (v[e] > 0.0f ? 1u : 0u) << e;
            ssign[i * ET + et] = (unsigned char)b;
        }
        FacLoads nxt = issue_fac(0, 0, 0);
        RowLoads rnx = issue_row(0);
        for (int t = 0; t < L; ++t) {
            const size_t tok = (size_t)t * R + rsafe;
            // head cotangents of this row (softmax backward of y_hat), lpg.py:83-84
            float dl[8], dpi;
            const RowLoads rc = rnx;
            if (t + 1 < L) rnx = issue_row(t + 1);
            {
                const float yh[8] = {rc.y0.x, rc.y0.y, rc.y0.z, rc.y0.w, rc.y1.x, rc.y1.y, rc.y1.z, rc.y1.w};
                const float dy[8] = {rc.d0.x, rc.d0.y, rc.d0.z, rc.d0.w, rc.d1.x, rc.d1.y, rc.d1.z, rc.d1.w};
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) s = fmaf(yh[i], dy[i], s);
#pragma unroll
                for (int i = 0; i < 8; ++i) dl[i] = rv ? S * (yh[i] * (dy[i] - s)) : 0.0f;
                dpi = rv ? S * rc.dpi : 0.0f;
                if (hf == 0 && rv) {
                    float4* qo = reinterpret_cast<float4*>(dl_out + tok * 8);
                    qo[0] = make_float4(dl[0], dl[1], dl[2], dl[3]);
                    qo[1] = make_float4(dl[4], dl[5], dl[6], dl[7]);
                }
            }
            // carry mask: the cell at step t-1 consumed (1 - done_{t-1}) * h_t
            const float nd = (t > 0 && rv && !rc.dn) ? 1.0f : 0.0f;
            const float nd_hp = rc.dn_hp ? 0.0f : 1.0f;
            if (t > 0) { mbar_wait(&q_full, (t - 1) & 1); tc_fence_after(); }
            const uint32_t p_addr = tmem_base + ((t - 1) & 1) * 256 + ((uint32_t)(q * 32) << 16);
            float dx3 = 0.f, dx4 = 0.f;
            for (int ub = 0; ub < 4; ++ub) {
                const int ubase = ub * 64 + hf * (64 / BT_NQ);       // first of this thread's 64 / BT_NQ units
#pragma unroll
                for (int c8 = 0; c8 < BT_NC; ++c8) {
                    const int u0 = ubase + c8 * 8;
                    float carry[8];
                    if (t > 0) tmem_ld8(p_addr + u0, carry);
                    const FacLoads cur = nxt;
                    {   // prefetch the next chunk (next c8, else next unit block, else next timestep)
                        int t2 = t, ub2 = ub, c2 = c8 + 1;
                        if (c2 == BT_NC) { c2 = 0; ++ub2; if (ub2 == 4) { ub2 = 0; ++t2; } }
                        if (t2 < L) nxt = issue_fac(t2, ub2, c2);
                    }
                    if (t > 0) tmem_ld_wait();
                    else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) carry[e] = 0.f;
                    }
                    float gr_[8], zz[8], gn_[8], hn_[8], hx[8];
                    unpack8h(cur.r, gr_); unpack8h(cur.z, zz); unpack8h(cur.n, gn_);
                    unpack8h(cur.hn, hn_); unpack8h(cur.hx, hx);
                    const int ci = ub * BT_NC + c8;
                    const unsigned sg = ssign[ci * ET + et];               // relu'(h_t)
                    {
                        unsigned b = 0;
#pragma unroll
                        for (int e = 0; e < 8; ++e) b |= (hx[e] > 0.0f ? 1u : 0u) << e;
                        ssign[ci * ET + et] = (unsigned char)b;           // relu'(h_{t+1}) for the next step
                    }
                    float gr[8], gz[8], ghn[8], gan[8], czh[8], czl[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int u = u0 + e;
                        float dh = nd * carry[e];
                        {
                            const float4 w0 = *reinterpret_cast<const float4*>(sWy + u * 8), w1 = *reinterpret_cast<const float4*>(sWy + u * 8 + 4);
                            float hd = dpi * swp[u];
                            hd = fmaf(dl[0], w0.x, hd); hd = fmaf(dl[1], w0.y, hd); hd = fmaf(dl[2], w0.z, hd); hd = fmaf(dl[3], w0.w, hd);
                            hd = fmaf(dl[4], w1.x, hd); hd = fmaf(dl[5], w1.y, hd); hd = fmaf(dl[6], w1.z, hd); hd = fmaf(dl[7], w1.w, hd);
                            dh += ((sg >> e) & 1u) ? hd : 0.0f;
                        }
                        if (!rv) dh = 0.0f;
                        // GRUCell backward factors from the saved gates (models/lpg.py:11-30):
                        //   d a_n = dh (1-z)(1-n^2);  d(Whn h + bhn) = d a_n * r;  d a_r = d a_n * hn * r(1-r);
                        //   d a_z = dh (h' - n) z(1-z)
                        const float omz = 1.0f - zz[e];
                        gan[e] = dh * omz * (1.0f - gn_[e] * gn_[e]);
                        ghn[e] = gan[e] * gr_[e];
                        gr[e] = ghn[e] * hn_[e] * (1.0f - gr_[e]);
                        gz[e] = dh * (nd_hp * hx[e] - gn_[e]) * zz[e] * omz;
                        const float cz = fminf(fmaxf(dh * zz[e], -65504.0f), 65504.0f);
                        czh[e] = __half2float(__float2half_rn(cz));
                        czl[e] = cz - czh[e];
                        {
                            const float4 w0 = *reinterpret_cast<const float4*>(sWi + u * 8), w1 = *reinterpret_cast<const float4*>(sWi + u * 8 + 4);
                            dx3 = fmaf(gr[e], w0.x, fmaf(gz[e], w0.y, fmaf(gan[e], w0.z, dx3)));
                            dx4 = fmaf(gr[e], w0.w, fmaf(gz[e], w1.x, fmaf(gan[e], w1.y, dx4)));
                        }
                    }
                    // the MMAs and the image stores of the previous unit block must have consumed the A stage
                    // (they finished long ago: this chunk's math alone takes longer)
                    if (c8 == 0 && ait > 0) mbar_wait(&a_empty, (ait & 1) ^ 1);
                    const uint32_t so = sA_u32 + sw128_offset(BT_M, rl, hf * (64 / BT_NQ) + c8 * 8);
                    st_shared_v4(so + 0 * BT_ACHUNK, pack8h_sat(gr));
                    st_shared_v4(so + 1 * BT_ACHUNK, pack8h_sat(gz));
                    st_shared_v4(so + 2 * BT_ACHUNK, pack8h_sat(ghn));
                    st_shared_v4(so + 3 * BT_ACHUNK, pack8h_sat(czh));
                    st_shared_v4(so + 4 * BT_ACHUNK, pack8h_sat(czl));
                    st_shared_v4(so + 5 * BT_ACHUNK, pack8h_sat(gan));
                    if (c8 & 1) {                                  // K-step hf * BT_NC / 2 + (c8 >> 1) of this unit block is complete
                        fence_proxy_async_smem();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&k_full[hf * (BT_NC / 2) + (c8 >> 1)]);
                    }
                }
                ++ait;
            }
            // d pyt / d pyt1: combine the unit slices of the row (fixed order)
            if (hf > 0) { sdx[((hf - 1) * BT_M + rl) * 2] = dx3; sdx[((hf - 1) * BT_M + rl) * 2 + 1] = dx4; }
            asm volatile("bar.sync 1, %0;" ::"n"(BT_EW * 32) : "memory");
            if (hf == 0 && rv) {
#pragma unroll
                for (int h2 = 0; h2 < BT_NQ - 1; ++h2) { dx3 += sdx[(h2 * BT_M + rl) * 2]; dx4 += sdx[(h2 * BT_M + rl) * 2 + 1]; }
                *reinterpret_cast<float2*>(dx + ((size_t)t * R + row) * 2) = make_float2(dx3, dx4);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(BT_EW * 32) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BT_EW + 1) tmem_dealloc(tmem_base, 512);
}

static size_t gru_bwd_tc_smem() {
    return BT_ASTAGE + BT_NSB * BT_BCHUNK + BT_ICHUNK + sizeof(float) * (LPG_H + LPG_H * LPG_Y + LPG_H * 8 + BT_M * 2 * (BT_NQ - 1)) +
           4 * BT_NC * BT_EW * 32 + 1024;
}

extern "C" int toued_gru_backward_tc(const uint8_t* done, const float* lpg_params, const void* whb_img,
                                     const void* h16, const void* fac, const float* y_hat, const float* d_pi_hat,
                                     const float* d_y_hat, void* dgimg, float* dl, float* dx, const uint32_t* cotangent_max,
                                     int n_agents, int n_workers, int rollout_len, int lifetime_conditioning, void* stream) {
    const int R = n_agents * n_workers;
    TOUED_CHECK(R > 0 && rollout_len > 0, "toued_gru_backward_tc: empty problem");
    const size_t smem = gru_bwd_tc_smem();
    TOUED_CUDA(cudaFuncSetAttribute(gru_backward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gru_backward_tc_kernel<<<(R + BT_M - 1) / BT_M, BT_THREADS, smem, (cudaStream_t)stream>>>(
        done, lpg_params, lifetime_conditioning ? 7 : 5, (const unsigned char*)whb_img, (const __half*)h16,
        (const __half*)fac, y_hat, d_pi_hat, d_y_hat, (unsigned char*)dgimg, dl, dx, cotangent_max, R, rollout_len, n_workers);
    TOUED_LAUNCH_CHECK();
    return 0;
}
