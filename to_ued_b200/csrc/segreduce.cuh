// Deterministic block-wide segmented reduction over an agent's row-sorted token list.
//
// The tabular agent kernels scatter per-token vectors (gradient / adjoint contributions) into table rows.
// With the tokens sorted by row (toued_sort_tokens) a row's tokens form one contiguous run, so the scatter
// is a segmented sum.  Popular rows (e.g. the reset state) have runs of several hundred tokens; summing a
// run sequentially in one thread serialises the whole CTA.  Here every thread sums a fixed chunk of
// consecutive sorted positions and the chunk partials are combined with a reverse segmented
// Hillis-Steele scan: the critical path is O(T/256 + log 256) and the summation tree is fixed, so results
// are bitwise reproducible (no atomics).  Block size must be 256.
#pragma once
#include "common.cuh"

struct SegIndex {
    uint16_t* tok;       // [T] token id at sorted position p
    uint16_t* row;       // [T] table row at sorted position p
    uint16_t* run;       // [T] run id of sorted position p
    uint16_t* run_row;   // [nruns] table row of run r
    int nruns;
    int chunk;           // positions per thread
};

__host__ __device__ inline size_t seg_index_bytes(int T) { return 4 * sizeof(uint16_t) * (size_t)T; }

// Builds the index in shared memory.  `mem` needs seg_index_bytes(T); `iscan` 2*256 ints.
__device__ inline SegIndex seg_index_build(void* mem, int* iscan, const uint16_t* __restrict__ sorted_tok,
                                           const int32_t* __restrict__ ob, int T) {
    SegIndex si;
    si.tok = reinterpret_cast<uint16_t*>(mem);
    si.row = si.tok + T;
    si.run = si.row + T;
    si.run_row = si.run + T;
    si.chunk = (T + 255) / 256;
    const int tid = threadIdx.x;
    for (int i = tid; i < T; i += 256) {
        const uint16_t tk = sorted_tok[i];
        si.tok[i] = tk;
        si.row[i] = (uint16_t)ob_idx(ob[tk]);
    }
    __syncthreads();
    const int p0 = tid * si.chunk, p1 = min(T, p0 + si.chunk);
    int cnt = 0;
    for (int p = p0; p < p1; ++p) cnt += (p == 0 || si.row[p - 1] != si.row[p]) ? 1 : 0;
    int* a = iscan; int* b = iscan + 256;
    a[tid] = cnt;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {                       // inclusive scan of the head counts
        b[tid] = a[tid] + (tid >= d ? a[tid - d] : 0);
        __syncthreads();
        int* t = a; a = b; b = t;
    }
    si.nruns = a[255];
    int r = a[tid] - cnt - 1;                                  // run id in front of this chunk
    for (int p = p0; p < p1; ++p) {
        if (p == 0 || si.row[p - 1] != si.row[p]) { ++r; si.run_row[r] = si.row[p]; }
        si.run[p] = (uint16_t)r;
    }
    __syncthreads();
    return si;
}

// runv[r][0..NV) = sum over the tokens of run r of rec[token][0..NV)   (rec row stride STRIDE floats).
// `scan` needs 2*256*NV floats, `flags` 2*256 bytes.  All 256 threads must call; ends with a barrier.
template <int NV, int STRIDE>
__device__ inline void seg_reduce(const float* __restrict__ rec, const SegIndex& si, int T,
                                  float* __restrict__ runv, float* scan, unsigned char* flags) {
    const int tid = threadIdx.x;
    const int p0 = tid * si.chunk, p1 = min(T, p0 + si.chunk);
    float acc[NV], first[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) { acc[j] = 0.f; first[j] = 0.f; }
    bool has_head = false;
    int cur = 0;
    for (int p = p0; p < p1; ++p) {
        if (p == 0 || si.row[p - 1] != si.row[p]) {
            if (has_head) {                                    // a run that starts and ends inside this chunk
#pragma unroll
                for (int j = 0; j < NV; ++j) runv[cur * NV + j] = acc[j];
            } else {
#pragma unroll
                for (int j = 0; j < NV; ++j) first[j] = acc[j];
            }
#pragma unroll
            for (int j = 0; j < NV; ++j) acc[j] = 0.f;
            has_head = true;
            cur = si.run[p];
        }
        const float* r = rec + (size_t)si.tok[p] * STRIDE;
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[j] += r[j];
    }
    // X = the part of this chunk that belongs to the run already open when the chunk began
    float* va = scan; float* vb = scan + 256 * NV;
    unsigned char* fa = flags; unsigned char* fb = flags + 256;
#pragma unroll
    for (int j = 0; j < NV; ++j) va[tid * NV + j] = has_head ? first[j] : acc[j];
    fa[tid] = has_head ? 1 : 0;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {                       // reverse segmented scan: T_m = X_m + (H_m ? 0 : T_{m+1})
        const bool take = (tid + d < 256) && !fa[tid];
#pragma unroll
        for (int j = 0; j < NV; ++j) vb[tid * NV + j] = va[tid * NV + j] + (take ? va[(tid + d) * NV + j] : 0.0f);
        fb[tid] = fa[tid] | (take ? fa[tid + d] : 0);
        __syncthreads();
        float* t = va; va = vb; vb = t;
        unsigned char* u = fa; fa = fb; fb = u;
    }
    if (has_head) {                                            // the run opened by this chunk's last head
#pragma unroll
        for (int j = 0; j < NV; ++j) runv[cur * NV + j] = acc[j] + (tid + 1 < 256 ? va[(tid + 1) * NV + j] : 0.0f);
    }
    __syncthreads();
}
