// "Row-chunk stream" kernel skeleton shared by the block-CSR SpMV and the level-scheduled
// ILU(0) sweeps (both are: for a set of rows, sum_j  A_ij * x_j  over short rows, then a small
// epilogue). Designed for HBM streaming on B200:
//
//   * a CTA owns a chunk of <= 256 consecutive stored rows whose entries fit a fixed shared-memory
//     tile (chunks are cut on the host, never across an ILU level);
//   * phase 1: the chunk's row offsets go to shared memory (one coalesced load);
//   * phase 2: ALL threads stream the chunk's column indices and value blocks with unit stride
//     (ld.global.cs: evict-first, they are touched once) and gather x[col] (default policy, so the
//     vector stays in L2); loads are issued U entries at a time per thread, so ~U*(8*bs^2+4+8*bs) B
//     are in flight per thread and the dependent chain is only  col -> x[col];
//   * phase 3: one thread per row adds its products from shared memory in ascending entry order
//     (the summation order of csr_mul_add!, src/StaticCSR/mat.jl:41-61) and runs the epilogue.
#pragma once
#include "jb_internal.cuh"
#include <cstdlib>

#define JB_CHUNK_ROWS 256
#define JB_CHUNK_CAP 2048   // entries (blocks) per chunk tile
#ifndef JB_STREAM_U
#define JB_STREAM_U 2      // entries in flight per thread in the streaming phase
#endif

template <int BS> struct StreamLoad;
template <> struct StreamLoad<1> {
    __device__ static __forceinline__ void mat(const double* __restrict__ v, size_t k, double (&a)[1]) { a[0] = __ldcs(v + k); }
    __device__ static __forceinline__ void vec(const double* x, int c, double (&o)[1]) { o[0] = x[c]; }
};
template <> struct StreamLoad<2> {
    __device__ static __forceinline__ void mat(const double* __restrict__ v, size_t k, double (&a)[4]) {
        const double2* p = reinterpret_cast<const double2*>(v + 4 * k);
        const double2 a0 = __ldcs(p), a1 = __ldcs(p + 1);
        a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y;
    }
    __device__ static __forceinline__ void vec(const double* x, int c, double (&o)[2]) {
        const double2 t = reinterpret_cast<const double2*>(x)[c];
        o[0] = t.x; o[1] = t.y;
    }
};
template <int BS> struct StreamLoadN {
    __device__ static __forceinline__ void mat(const double* __restrict__ v, size_t k, double (&a)[BS * BS]) {
#pragma unroll
        for (int i = 0; i < BS * BS; i++) a[i] = __ldcs(v + k * BS * BS + i);
    }
    __device__ static __forceinline__ void vec(const double* x, int c, double (&o)[BS]) {
#pragma unroll
        for (int i = 0; i < BS; i++) o[i] = x[(size_t)c * BS + i];
    }
};
template <> struct StreamLoad<3> : StreamLoadN<3> {};
template <> struct StreamLoad<4> : StreamLoadN<4> {};

// Phases 1-2 for one chunk. On return (after the trailing __syncthreads) s_rp[0..nr] holds the chunk's row
// offsets and s_prod[e*BS + q] the product of entry (base + e) with its x block.
template <int BS, int U>
__device__ __forceinline__ void stream_chunk_products(int t0, int nr, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                      const double* __restrict__ val, size_t val_block_offset, const double* x,
                                                      int32_t* s_rp, double* s_prod) {
    for (int j = threadIdx.x; j <= nr; j += blockDim.x) s_rp[j] = __ldg(rowptr + t0 + j);
    __syncthreads();
    const int base = s_rp[0];
    const int cnt = s_rp[nr] - base;
    for (int e0 = threadIdx.x; e0 < cnt; e0 += blockDim.x * U) {
        int col[U];
        double a[U][BS * BS];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = e0 + u * blockDim.x;
            if (e < cnt) {
                col[u] = __ldcs(colidx + base + e);
                StreamLoad<BS>::mat(val, val_block_offset + (size_t)(base + e), a[u]);
            }
        }
        double xv[U][BS];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = e0 + u * blockDim.x;
            if (e < cnt) StreamLoad<BS>::vec(x, col[u], xv[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = e0 + u * blockDim.x;
            if (e < cnt) {
#pragma unroll
                for (int i = 0; i < BS; i++) {
                    double s = a[u][i] * xv[u][0];
#pragma unroll
                    for (int p = 1; p < BS; p++) s = fma(a[u][p * BS + i], xv[u][p], s);
                    s_prod[(size_t)e * BS + i] = s;
                }
            }
        }
    }
    __syncthreads();
}

// Phase 3 helper: sum of the products of local row j (ascending entry order).
template <int BS>
__device__ __forceinline__ void stream_row_sum(int j, const int32_t* s_rp, const double* s_prod, double (&acc)[BS]) {
    const int base = s_rp[0];
#pragma unroll
    for (int i = 0; i < BS; i++) acc[i] = 0.0;
    for (int e = s_rp[j] - base; e < s_rp[j + 1] - base; e++) {
#pragma unroll
        for (int i = 0; i < BS; i++) acc[i] += s_prod[(size_t)e * BS + i];
    }
}

// Host: cut rows [r0, r1) (offsets `ptr`, indexed by stored row) into chunks of <= JB_CHUNK_ROWS rows and
// <= JB_CHUNK_CAP entries; appends chunk start rows to `chunks` (caller appends the final end).
// Returns false if a single row exceeds the tile.
// Rows per chunk: JB_CHUNK_ROWS. Smaller chunks for small row sets (to shorten the tail of the persistent grid) were
// measured and lose: at 1.26M rows, 64..96-row chunks make the ILU sweeps 44 % and the SpMV 22 % slower than 256-row
// chunks (per-chunk barriers and row-offset loads dominate). The environment variable JB_CHUNK_ROWS (32..256) is kept
// for experiments.
inline int32_t jb_pick_chunk_rows(int64_t nrows) {
    (void)nrows;
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("JB_CHUNK_ROWS"); forced = e ? atoi(e) : 0; }
    if (forced >= 32 && forced <= JB_CHUNK_ROWS) return forced;
    return JB_CHUNK_ROWS;
}
inline bool jb_cut_chunks(const std::vector<int32_t>& ptr, int32_t r0, int32_t r1, std::vector<int32_t>& chunks) {
    const int32_t max_rows = jb_pick_chunk_rows((int64_t)r1 - r0);
    int32_t start = r0;
    while (start < r1) {
        int32_t end = start;
        while (end < r1 && end - start < max_rows && ptr[end + 1] - ptr[start] <= JB_CHUNK_CAP) end++;
        if (end == start) return false;
        chunks.push_back(start);
        start = end;
    }
    return true;
}
