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

constexpr int WT_THREADS = 192;            // 4 epilogue warps + producer warp + MMA warp
constexpr int WT_NS = 3;
constexpr int WT_STAGE = 8 * 8192;         // 2 A sub-tiles + 6 B sub-tiles of 8 KB

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_wh_tc_kernel(const unsigned char* __restrict__ hpimg, const unsigned char* __restrict__ dgimg,
                   float* __restrict__ partial, int n_tok_blocks, int blocks_per_split, int accumulate) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full[WT_NS], empty[WT_NS], done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int type = blockIdx.x & 3, split = blockIdx.x >> 2;
    const int jt = type >> 1, ct = type & 1;
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
                mbar_expect_tx(&full[s], WT_STAGE);
                unsigned char* st = smem + s * WT_STAGE;
                const size_t tb = (size_t)(tb0 + i);
                bulk_g2s(st, hpimg + ((tb * 4 + 2 * jt) << 13), 2 * 8192, &full[s]);            // j groups 2jt, 2jt+1
                bulk_g2s(st + 2 * 8192, dgimg + ((tb * 16 + 6 * ct) << 13), 6 * 8192, &full[s]);  // c groups 6ct .. 6ct+5
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc_mn(128, 192, 1);
            for (int i = 0; i < nblk; ++i) {
                const int s = i % WT_NS;
                mbar_wait(&full[s], (i / WT_NS) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * WT_STAGE), b0 = a0 + 2 * 8192;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t ad = tc_smem_desc_mn(a0 + ks * 2048, 8192);
                    tc_mma(tmem_base, ad, tc_smem_desc_mn(b0 + ks * 2048, 8192), idesc, (i | ks) != 0);
                    tc_mma(tmem_base + 192, ad, tc_smem_desc_mn(b0 + 3 * 8192 + ks * 2048, 8192), idesc, (i | ks) != 0);
                }
                tc_commit(&empty[s]);
            }
            tc_commit(&done_bar);
        }
    } else {
        // epilogue: TMEM lane = j within the tile
        mbar_wait(&done_bar, 0);
        tc_fence_after();
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
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// small gradients from the bf16 dG image (TC path).  Same partial layout as wgrad_small_kernel.
constexpr int SMT_WI = 0, SMT_BHN = 8 * LPG_G, SMT_WPI = SMT_BHN + LPG_H, SMT_WY = SMT_WPI + LPG_H,
              SMT_BPI = SMT_WY + LPG_H * LPG_Y, SMT_BY = SMT_BPI + 1, SMT_TOTAL = SMT_BY + LPG_Y;

__global__ void __launch_bounds__(256)
wgrad_small_tc_kernel(const float* __restrict__ x, const __half* __restrict__ h16, const unsigned char* __restrict__ dgimg,
                      const float* __restrict__ d_pi_hat, const float* __restrict__ dl, float* __restrict__ partial,
                      int R, int L, int toks_per_split, int accumulate) {
    const int j = threadIdx.x, split = blockIdx.x;
    const size_t ntok = (size_t)L * R;
    const size_t Rp = ((size_t)R + 63) & ~(size_t)63;
    const size_t t0 = (size_t)split * toks_per_split;
    const size_t t1 = min(ntok, t0 + (size_t)toks_per_split);
    float wi[3][8], bhn = 0.f, head[9], hb = 0.f;
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int q = 0; q < 8; ++q) wi[g][q] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) head[i] = 0.f;
    const int cgj = j >> 6, cin = j & 63;
    for (size_t tok = t0; tok < t1; ++tok) {
        const size_t t = tok / R, row = tok % R;
        const size_t itok = t * Rp + row;
        const float4 x0 = *reinterpret_cast<const float4*>(x + tok * 8), x1 = *reinterpret_cast<const float4*>(x + tok * 8 + 4);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const float4 d0 = *reinterpret_cast<const float4*>(dl + tok * 8), d1 = *reinterpret_cast<const float4*>(dl + tok * 8 + 4);
        const float dv[9] = {d_pi_hat[tok], d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        auto ld = [&](int gate) {
            return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(dgimg + tile_img_offset(itok, 16, (gate * 4 + cgj) * 64 + cin)));
        };
        const float dar = ld(0), daz = ld(1), dhn = ld(2), dan = ld(3);
        const float y = fmaxf(__half2float(h16[tok * LPG_H + j]), 0.0f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            wi[0][q] = fmaf(xv[q], dar, wi[0][q]); wi[1][q] = fmaf(xv[q], daz, wi[1][q]); wi[2][q] = fmaf(xv[q], dan, wi[2][q]);
        }
        bhn += dhn;
#pragma unroll
        for (int i = 0; i < 9; ++i) head[i] = fmaf(y, dv[i], head[i]);
        if (j < 9) hb += dv[j];
    }
    float* out = partial + (size_t)split * SMT_TOTAL;
    auto put = [&](int idx, float v) { out[idx] = accumulate ? out[idx] + v : v; };
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int q = 0; q < 8; ++q) put(SMT_WI + q * LPG_G + g * LPG_H + j, wi[g][q]);
    put(SMT_BHN + j, bhn);
    put(SMT_WPI + j, head[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) put(SMT_WY + j * 8 + i, head[1 + i]);
    if (j < 9) put(SMT_BPI + j, hb);
}

constexpr int WT_SPLITS = 37;              // 4 tile types x 37 = 148 CTAs
constexpr int SMT_SPLITS = 592;

extern "C" int toued_wgrad_tc_splits(void) { return WT_SPLITS; }

extern "C" int toued_lpg_wgrad_tc(const void* hpimg, const void* dgimg, const float* x, const void* h16,
                                  const float* d_pi_hat, const float* dl, float* wh_partials, float* small_partials,
                                  int n_agents, int n_workers, int rollout_len, int accumulate, void* stream) {
    const int R = n_agents * n_workers, L = rollout_len;
    TOUED_CHECK(R > 0 && L > 0, "toued_lpg_wgrad_tc: empty problem");
    cudaStream_t st = (cudaStream_t)stream;
    const int Rp = (R + 63) / 64 * 64;
    const int n_tb = L * Rp / 64;
    const int bps = (n_tb + WT_SPLITS - 1) / WT_SPLITS;
    const size_t smem = WT_NS * WT_STAGE + 1024;
    TOUED_CUDA(cudaFuncSetAttribute(wgrad_wh_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_wh_tc_kernel<<<4 * WT_SPLITS, WT_THREADS, smem, st>>>((const unsigned char*)hpimg, (const unsigned char*)dgimg,
                                                                wh_partials, n_tb, bps, accumulate);
    TOUED_LAUNCH_CHECK();
    const size_t ntok = (size_t)L * R;
    const int tps = (int)((ntok + SMT_SPLITS - 1) / SMT_SPLITS);
    wgrad_small_tc_kernel<<<SMT_SPLITS, 256, 0, st>>>(x, (const __half*)h16, (const unsigned char*)dgimg, d_pi_hat, dl,
                                                      small_partials, R, L, tps, accumulate);
    TOUED_LAUNCH_CHECK();
    return 0;
}
