// "TMA-staged row stream" skeleton for 2x2-block CSR data (SpMV and the level-scheduled ILU(0) sweeps), sm_100a.
//
// Second generation of jb_stream.cuh, built like the assembly kernel (assembly.cu, twophase_assemble_pair_kernel):
//   * a persistent CTA (4 warps) walks a host-built chunk table (<= 64 stored rows / <= 448 blocks per chunk);
//   * ONE lane issues 1-D bulk copies (cp.async.bulk + mbarrier, jb_tma.cuh) of the chunk's row offsets, column
//     indices and 32-byte value blocks into one of two shared-memory stages; the copies of chunk i+1 are in flight
//     while chunk i is consumed, so the matrix stream never waits on a thread and holds no registers;
//   * lane pair (2j, 2j+1) of a warp owns stored row j of the warp's 16 rows: lane e accumulates component e of
//     sum_k A_jk x_k over the row's blocks in ascending order (the order of csr_mul_add!, src/StaticCSR/mat.jl:41-61,
//     and of forward/backward_substitute!, src/StaticCSR/ilu0.jl:156-187) entirely in registers; the x gathers of up
//     to NG blocks are issued before the first product;
//   * no shared-memory partial products and no CTA barrier in steady state: a stage is refilled by the last warp to
//     finish with it (shared-memory counter).
// The caller supplies the epilogue (SpMV: y, fused inner products; sweep: rhs - sum, optional D^{-1}).
#pragma once
#include <algorithm>
#include <vector>

#include "jb_internal.cuh"
#include "jb_tma.cuh"

#define JB_S2_ROWS 64
#define JB_S2_CAP 448
#define JB_S2_THREADS 128
#define JB_S2_IDENT_ROWS 512   // rows per identity super-chunk of the SpMV table (krylov.cu)
#define JB_S2_CTAS_PER_SM 6    // resident CTAs per SM of the s2 kernels (launch bounds): grid of the persistent kernels = SMs x 6

struct __align__(16) S2Stage {
    double val[JB_S2_CAP * 4];        // 32-byte blocks: always aligned, no lead
    int32_t col[JB_S2_CAP + 4];
    int32_t rp[JB_S2_ROWS + 8];
    int32_t ord[JB_S2_ROWS + 4];      // row ids of the stored rows (sweeps; unused by the SpMV)
};
static_assert(offsetof(S2Stage, col) % 16 == 0 && offsetof(S2Stage, rp) % 16 == 0 && offsetof(S2Stage, ord) % 16 == 0 &&
              sizeof(S2Stage) % 16 == 0, "stage alignment");

struct S2Smem {
    S2Stage stage[2];
    S2Chunk meta[2];
    uint64_t bar[2];
    int cnt[2];
};

// Issue the copies of table entry k into stage s (one lane). val_block_offset: first block of this matrix inside `val`
// (the ILU value buffer is [L | D | U]); order == nullptr: row ids are t0 + j.
__device__ __forceinline__ void s2_issue(S2Smem& sm, int s, const S2Chunk* __restrict__ table, int k, const int32_t* __restrict__ ptr,
                                         const int32_t* __restrict__ col, const double* __restrict__ val, size_t val_block_offset,
                                         const int32_t* __restrict__ order) {
    const int4* tp = reinterpret_cast<const int4*>(table + k);
    const int4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
    S2Chunk ci;
    ci.t0 = t0.x; ci.nr = t0.y; ci.e0 = t0.z; ci.cnt = t0.w; ci.flags = t1.x; ci.pad0 = ci.pad1 = ci.pad2 = 0;
    sm.meta[s] = ci;
    if (ci.flags & 1) {   // identity chunk: nothing to stream
        jb_mbar_expect_tx(&sm.bar[s], 0);
        return;
    }
    S2Stage& S = sm.stage[s];
    const JbSpan<int32_t> sC(col, (size_t)ci.e0, ci.cnt), sP(ptr, (size_t)ci.t0, ci.nr + 1);
    const uint32_t bV = (uint32_t)ci.cnt * 32u;
    uint32_t total = sC.bytes + sP.bytes + bV;
    uint32_t bO = 0;
    const char* oSrc = nullptr;
    if (order) { const JbSpan<int32_t> sO(order, (size_t)ci.t0, ci.nr); bO = sO.bytes; oSrc = sO.src; total += bO; }
    jb_mbar_expect_tx(&sm.bar[s], total);
    const uint64_t pol = jb_policy_evict_first();
    if (sP.bytes) jb_bulk_g2s_hint(S.rp, sP.src, sP.bytes, &sm.bar[s], pol);
    if (sC.bytes) jb_bulk_g2s_hint(S.col, sC.src, sC.bytes, &sm.bar[s], pol);
    if (bO) jb_bulk_g2s_hint(S.ord, oSrc, bO, &sm.bar[s], pol);
    if (bV) jb_bulk_g2s_hint(S.val, val + (val_block_offset + (size_t)ci.e0) * 4, bV, &sm.bar[s], pol);
}

__device__ __forceinline__ void s2_prologue(S2Smem& sm, const S2Chunk* __restrict__ table, int c0, int c1, const int32_t* __restrict__ ptr,
                                            const int32_t* __restrict__ col, const double* __restrict__ val, size_t val_block_offset,
                                            const int32_t* __restrict__ order) {
    if (threadIdx.x == 0) {
        jb_mbar_init(&sm.bar[0], 1);
        jb_mbar_init(&sm.bar[1], 1);
        sm.cnt[0] = 0; sm.cnt[1] = 0;
        jb_mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int k0 = c0 + (int)blockIdx.x, k1 = k0 + (int)gridDim.x;
        if (k0 < c1) s2_issue(sm, 0, table, k0, ptr, col, val, val_block_offset, order);
        if (k1 < c1) s2_issue(sm, 1, table, k1, ptr, col, val, val_block_offset, order);
    }
    __syncthreads();
}

// Hand stage s back: the last of the CTA's warps to arrive refills it with the chunk two grid strides ahead.
__device__ __forceinline__ void s2_release(S2Smem& sm, int s, const S2Chunk* __restrict__ table, int k, int c1, const int32_t* __restrict__ ptr,
                                           const int32_t* __restrict__ col, const double* __restrict__ val, size_t val_block_offset,
                                           const int32_t* __restrict__ order) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        __threadfence_block();
        const int old = atomicAdd(&sm.cnt[s], 1);
        if (old == JB_S2_THREADS / 32 - 1) {
            sm.cnt[s] = 0;
            const int k2 = k + 2 * (int)gridDim.x;
            if (k2 < c1) s2_issue(sm, s, table, k2, ptr, col, val, val_block_offset, order);
        }
    }
}

// Component e (= lane & 1) of sum_k A_jk x_k for the stored row j of this lane pair; blocks k0..k1 (stage-local).
template <int NG>
__device__ __forceinline__ double s2_row_sum(const S2Stage& S, int lead_col, int k0, int k1, int e, const double* x) {
    double acc = 0.0;
    for (int kb = k0; kb < k1; kb += NG) {
        double2 xv[NG];
#pragma unroll
        for (int u = 0; u < NG; u++)
            if (kb + u < k1) xv[u] = reinterpret_cast<const double2*>(x)[S.col[lead_col + kb + u]];
#pragma unroll
        for (int u = 0; u < NG; u++)
            if (kb + u < k1) {
                const double* a = S.val + 4 * (size_t)(kb + u);
                double sacc = a[e] * xv[u].x;
                sacc = fma(a[2 + e], xv[u].y, sacc);
                acc += sacc;
            }
    }
    return acc;
}

// Host: cut stored rows [r0, r1) (offsets ptr) into chunks; returns false if a row exceeds the stage.
// balance_ctas > 0: the persistent kernels deal chunk k to CTA k mod grid, so with n chunks of 64 rows the busiest CTA
// streams ceil(n / grid) * 64 rows while the average is n * 64 / grid — 8 % more at 11.1 chunks per CTA (1.26M cells per
// GPU). The rows are therefore cut into a multiple of `balance_ctas` chunks of equal size (<= 64 rows): every CTA gets the
// same number of rows. Large row sets (>= 64 chunks per CTA, quantisation < 1.6 %) keep full 64-row chunks.
inline bool jb_s2_cut(const std::vector<int32_t>& ptr, int32_t r0, int32_t r1, std::vector<S2Chunk>& out, int balance_ctas = 0) {
    int32_t start = r0;
    int64_t target = 0;     // number of chunks aimed at (0: greedy 64-row chunks)
    const int64_t rows = (int64_t)r1 - r0;
    if (balance_ctas > 0 && rows > 0) {
        const int64_t full = (rows + JB_S2_ROWS - 1) / JB_S2_ROWS;
        const int64_t per_cta = (full + balance_ctas - 1) / balance_ctas;
        if (per_cta < 64) target = per_cta * balance_ctas;
    }
    int64_t made = 0;
    while (start < r1) {
        int32_t limit = JB_S2_ROWS;
        if (target > made) {      // spread the remaining rows evenly over the remaining chunks of the target
            const int64_t left = (int64_t)r1 - start, chunks_left = target - made;
            limit = (int32_t)std::min<int64_t>(JB_S2_ROWS, std::max<int64_t>(1, (left + chunks_left - 1) / chunks_left));
        }
        int32_t end = start;
        while (end < r1 && end - start < limit && ptr[end + 1] - ptr[start] <= JB_S2_CAP) end++;
        made++;
        if (end == start) return false;
        S2Chunk c;
        c.t0 = start; c.nr = end - start; c.e0 = ptr[start]; c.cnt = ptr[end] - ptr[start]; c.flags = 0; c.pad0 = c.pad1 = c.pad2 = 0;
        out.push_back(c);
        start = end;
    }
    return true;
}
