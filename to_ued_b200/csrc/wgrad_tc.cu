// Tensor-core weight gradient of the recurrent matrix:  dWh[j][c] = sum_tok h'[tok][j] * dGh[tok][c]
// (the parameter-gradient sum of jax.grad through models/lpg.py:29, flax GRUCell hr/hz/hn kernels).
//
// Both operands are contracted over tokens, i.e. they are MN-major for tcgen05: the 64-token x 64-column
// sub-tiles of the bf16 token tile images written by the forward (h') and backward (dG) kernels are
// bulk-copied (cp.async.bulk) into shared memory and used as-is.  A CTA owns one 128 (j) x 384 (c)
// output tile (fp32 accumulators in 384 TMEM columns) and a contiguous range of token blocks; the four
// tile types of one token range are adjacent CTAs so they share their operand stream through L2.
// Output: per-split partial sums (deterministic; reduced by toued_reduce_partials).
#include "tc.cuh"
#include "lpg_common.cuh"
#include "../../include/toued.h"

// small-gradient partial layout (identical to wgrad_small_kernel in lpg_backward.cu)
constexpr int SMT_WI = 0, SMT_BHN = 8 * LPG_G, SMT_WPI = SMT_BHN + LPG_H, SMT_WY = SMT_WPI + LPG_H,
              SMT_BPI = SMT_WY + LPG_H * LPG_Y, SMT_BY = SMT_BPI + 1, SMT_TOTAL = SMT_BY + LPG_Y;
constexpr int WT_THREADS = 192;            // 4 epilogue warps + producer warp + MMA warp
constexpr int WT_NS = 3;
constexpr int WT_STAGE = 9 * 8192;         // dWh tiles: 2 A + 6 B sub-tiles of 8 KB; x tiles: 1 A + 8 B

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_wh_tc_kernel(const unsigned char* __restrict__ hpimg, const unsigned char* __restrict__ dgimg,
                   const unsigned char* __restrict__ ximg, float* __restrict__ partial, float* __restrict__ small_partial,
                   int n_tok_blocks, int blocks_per_split, int accumulate) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full[WT_NS], empty[WT_NS], done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tile types 0-3: dWh tile (jt, ct) = h'^T [128 j] x dG [384 c];
    // tile types 4-5: input-side gradients  x^T [8 (of 64) inputs] x dG column groups 8*(type-4) .. +8
    const int type = blockIdx.x % 6, split = blockIdx.x / 6;
    const bool xt = type >= 4;
    const int jt = (type >> 1) & 1, ct = type & 1;
    const int tb0 = split * blocks_per_split;
    const int tb1 = min(n_tok_blocks, tb0 + blocks_per_split);
    const int nblk = max(0, tb1 - tb0);

    if (tid == 0) {
        for (int s = 0; s < WT_NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 4) {
        if (lane == 0) {
            for (int i = 0; i < nblk; ++i) {
                const int s = i % WT_NS;
                mbar_wait(&empty[s], ((i / WT_NS) & 1) ^ 1);
                mbar_expect_tx(&full[s], xt ? 9 * 8192 : 8 * 8192);
                unsigned char* st = smem + s * WT_STAGE;
                const size_t tb = (size_t)(tb0 + i);
                if (!xt) {
                    bulk_g2s(st, hpimg + ((tb * 4 + 2 * jt) << 13), 2 * 8192, &full[s]);            // j groups 2jt, 2jt+1
                    bulk_g2s(st + 2 * 8192, dgimg + ((tb * 16 + 6 * ct) << 13), 6 * 8192, &full[s]);  // c groups 6ct .. 6ct+5
                } else {
                    bulk_g2s(st, ximg + (tb << 13), 8192, &full[s]);                                  // the x group
                    bulk_g2s(st + 8192, dgimg + ((tb * 16 + 8 * ct) << 13), 7 * 8192, &full[s]);      // c groups 8ct .. 8ct+6
                    bulk_g2s(st + 8 * 8192, dgimg + ((tb * 16 + 8 * ct + 7) << 13), 8192, &full[s]);  // c group 8ct+7
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc_mn(128, 192, 1), idesc_x = tc_idesc_mn(128, 256, 1);
            for (int i = 0; i < nblk; ++i) {
                const int s = i % WT_NS;
                mbar_wait(&full[s], (i / WT_NS) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * WT_STAGE);
                if (!xt) {
                    const uint32_t b0 = a0 + 2 * 8192;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = tc_smem_desc_mn(a0 + ks * 2048, 8192);
                        tc_mma(tmem_base, ad, tc_smem_desc_mn(b0 + ks * 2048, 8192), idesc, (i | ks) != 0);
                        tc_mma(tmem_base + 192, ad, tc_smem_desc_mn(b0 + 3 * 8192 + ks * 2048, 8192), idesc, (i | ks) != 0);
                    }
                } else {
                    const uint32_t b0 = a0 + 8192;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        // LBO = 0: both 64-row halves of the M = 128 operand alias the single x group
                        const uint64_t ad = tc_smem_desc_mn(a0 + ks * 2048, 0);
                        tc_mma(tmem_base, ad, tc_smem_desc_mn(b0 + ks * 2048, 8192), idesc_x, (i | ks) != 0);
                        tc_mma(tmem_base + 256, ad, tc_smem_desc_mn(b0 + 4 * 8192 + ks * 2048, 8192), idesc_x, (i | ks) != 0);
                    }
                }
                tc_commit(&empty[s]);
            }
            tc_commit(&done_bar);
        }
    } else {
        // epilogue: TMEM lane = j within the tile
        mbar_wait(&done_bar, 0);
        tc_fence_after();
        if (xt) {
            // rows 0..7 of the x tile: dWi (rows q < X), dbi (row 7) and dbhn (row 7 of the dhn columns).
            // Always accumulates: the head-gradient kernel has initialised this split's small-partial area.
            if (warp == 0) {
                float* out = small_partial + (size_t)split * SMT_TOTAL;
                for (int c = 0; c < 512; c += 8) {
                    float v[8];
                    if (nblk > 0) { tmem_ld8(tmem_base + c, v); tmem_ld_wait(); }
                    else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = 0.f;
                    }
                    if (lane < 8) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int cc = c + e;
                            int dst = -1;
                            if (ct == 0) dst = SMT_WI + lane * LPG_G + cc;                       // dar, daz columns
                            else if (cc >= 256) dst = SMT_WI + lane * LPG_G + 512 + (cc - 256);  // dan columns
                            else if (lane == 7) dst = SMT_BHN + cc;                              // sum of dhn
                            if (dst >= 0) out[dst] += v[e];
                        }
                    }
                }
            }
        } else {
        const int j = jt * 128 + warp * 32 + lane;
        float* out = partial + (size_t)split * LPG_H * LPG_G + (size_t)j * LPG_G + ct * 384;
        for (int c = 0; c < 384; c += 8) {
            float v[8];
            if (nblk > 0) {
                tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;
            }
            float4* p = reinterpret_cast<float4*>(out + c);
            if (accumulate) {
                const float4 p0 = p[0], p1 = p[1];
                v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
            }
            p[0] = make_float4(v[0], v[1], v[2], v[3]);
            p[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// head gradients (dw_pi, dW_y, db_pi, db_y) + initialisation of the split's input-side area (TC path).
// A warp owns one 8-unit chunk of the RB32 layout: lane = row of a 32-row block, so every h16 access is a
// coalesced 512-byte line; each lane keeps 8 x 9 accumulators over all the row blocks of its split and the
// 32 rows are combined once at the end by a fixed shuffle tree (deterministic).
constexpr int HT_SPLITS = 148;
__global__ void __launch_bounds__(256)
wgrad_heads_tc_kernel(const __half* __restrict__ h16, const float* __restrict__ d_pi_hat, const float* __restrict__ dl,
                      float* __restrict__ partial, int R, int L, int blocks_per_split, int accumulate) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * 8 + warp;            // 0..31
    const int split = blockIdx.y;
    const int R32 = (R + 31) >> 5;
    const int nrb = L * R32;                             // (t, row block) pairs
    const int b0 = split * blocks_per_split, b1 = min(nrb, b0 + blocks_per_split);
    float acc[8][9], hb[9];
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[e][i] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) hb[i] = 0.f;
    for (int b = b0; b < b1; ++b) {
        const int t = b / R32, rb = b % R32;
        const int row = rb * 32 + lane;
        if (row >= R) continue;
        const size_t tok = (size_t)t * R + row;
        const uint4 raw = *reinterpret_cast<const uint4*>(h16 + (((size_t)b * 32 + chunk) << 8) + (lane << 3));
        const float4 d0 = *reinterpret_cast<const float4*>(dl + tok * 8), d1 = *reinterpret_cast<const float4*>(dl + tok * 8 + 4);
        const float dv[9] = {d_pi_hat[tok], d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const __half2* hh = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
            const float2 f = __half22float2(hh[e2]);
            const float y0 = fmaxf(f.x, 0.0f), y1 = fmaxf(f.y, 0.0f);
#pragma unroll
            for (int i = 0; i < 9; ++i) { acc[2 * e2][i] = fmaf(y0, dv[i], acc[2 * e2][i]); acc[2 * e2 + 1][i] = fmaf(y1, dv[i], acc[2 * e2 + 1][i]); }
        }
        if (chunk == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) hb[i] += dv[i];
        }
    }
    float* out = partial + (size_t)split * SMT_TOTAL;
    auto put = [&](int idx, float v) { out[idx] = accumulate ? out[idx] + v : v; };
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float v = acc[e][i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) { if (i == 0) put(SMT_WPI + chunk * 8 + e, v); else put(SMT_WY + (chunk * 8 + e) * 8 + (i - 1), v); }
        }
    if (chunk == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float v = hb[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) put(SMT_BPI + i, v);
        }
    }
    if (!accumulate) {                    // the x-tile CTAs of the GEMM kernel add into the input-side area
        for (int i = blockIdx.x * 256 + threadIdx.x; i < SMT_WPI; i += 4 * 256) out[i] = 0.0f;
    }
}

constexpr int WT_SPLITS = 24;              // 6 tile types x 24 = 144 CTAs
constexpr int SMT_SPLITS = 592;

extern "C" int toued_wgrad_tc_splits(void) { return WT_SPLITS; }

extern "C" int toued_lpg_wgrad_tc(const void* hpimg, const void* dgimg, const void* ximg, const void* h16,
                                  const float* d_pi_hat, const float* dl, float* wh_partials, float* small_partials,
                                  int n_agents, int n_workers, int rollout_len, int accumulate, void* stream) {
    const int R = n_agents * n_workers, L = rollout_len;
    TOUED_CHECK(R > 0 && L > 0, "toued_lpg_wgrad_tc: empty problem");
    cudaStream_t st = (cudaStream_t)stream;
    const int Rp = (R + 63) / 64 * 64;
    const int n_tb = L * Rp / 64;
    const int bps = (n_tb + WT_SPLITS - 1) / WT_SPLITS;
    // head gradients first: also zero-initialises the input-side area the GEMM's x tiles accumulate into
    const int nrb = L * ((R + 31) / 32);
    const int rbps = (nrb + HT_SPLITS - 1) / HT_SPLITS;
    wgrad_heads_tc_kernel<<<dim3(4, HT_SPLITS), 256, 0, st>>>((const __half*)h16, d_pi_hat, dl, small_partials, R, L, rbps, accumulate);
    TOUED_LAUNCH_CHECK();
    const size_t smem = WT_NS * WT_STAGE + 1024;
    TOUED_CUDA(cudaFuncSetAttribute(wgrad_wh_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_wh_tc_kernel<<<6 * WT_SPLITS, WT_THREADS, smem, st>>>((const unsigned char*)hpimg, (const unsigned char*)dgimg,
                                                                (const unsigned char*)ximg, wh_partials, small_partials,
                                                                n_tb, bps, accumulate);
    TOUED_LAUNCH_CHECK();
    return 0;
}
