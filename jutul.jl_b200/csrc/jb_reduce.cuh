// Deterministic grid-wide reductions: every CTA writes its partial sums, the last
// CTA to finish (ticket counter) adds the partials in a fixed order and runs a
// finaliser with the totals. No floating-point atomics, so results are bitwise
// reproducible for a fixed grid size.
#pragma once
#include "jb_internal.cuh"

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Warp-aggregated FP64 atomic add for the face-parallel scatters (fvm_face_assembly!, src/conservation/fvm_assembly.jl:253-283):
// lanes of the calling warp that target the same address (faces listed by cell: neighbouring lanes share a cell) are found
// with match.any, their values are added in lane order by every member and ONE lane issues the atomic. Correct for any
// subset of converged lanes; cuts the atomic traffic of a face sweep by the number of faces a cell owns.
__device__ __forceinline__ void warp_agg_atomic_add(double* addr, double v) {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, (unsigned long long)addr);
    double sum = 0.0;
    for (unsigned rem = peers; rem; rem &= rem - 1) sum += __shfl_sync(peers, v, __ffs(rem) - 1);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(addr, sum);
}
// NaN-propagating max: a non-finite residual must reach the host (fmax would drop the NaN)
__device__ __forceinline__ double nan_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : fmax(a, b)); }
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct OpSum { __device__ static double ident() { return 0.0; } __device__ static double warp(double v) { return warp_sum(v); } __device__ static double comb(double a, double b) { return a + b; } };
struct OpMax { __device__ static double ident() { return 0.0; } __device__ static double warp(double v) { return warp_max(v); } __device__ static double comb(double a, double b) { return nan_max(a, b); } };

// Reduce NR values over the CTA; result valid in thread 0. BLOCK must be a multiple of 32, <= 1024.
template <int NR, class Op>
__device__ __forceinline__ void block_reduce(double (&v)[NR], double* smem /* NR*32 */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int r = 0; r < NR; r++) v[r] = Op::warp(v[r]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < NR; r++) smem[r * 32 + wid] = v[r];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
            double x = (lane < nw) ? smem[r * 32 + lane] : Op::ident();
            v[r] = Op::warp(x);
        }
    }
}

// Grid reduction. All threads of all CTAs must call. `fin(totals)` runs on thread 0 of the
// last CTA. partials: gridDim.x*NR doubles; counter: one unsigned, zero before first use
// (atomicInc wraps it back to zero).
template <int NR, class Op, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NR], double* partials, unsigned int* counter, Fin fin) {
    __shared__ double smem[NR * 32];
    __shared__ bool is_last;
    block_reduce<NR, Op>(v, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NR; r++) partials[(size_t)blockIdx.x * NR + r] = v[r];
        __threadfence();
        unsigned int ticket = atomicInc(counter, gridDim.x - 1);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) acc[r] = Op::ident();
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int r = 0; r < NR; r++) acc[r] = Op::comb(acc[r], __ldcg(&partials[(size_t)b * NR + r]));
        }
        block_reduce<NR, Op>(acc, smem);
        if (threadIdx.x == 0) fin(acc);
    }
}

// Two-launch form: kernel A (finalize = false) only deposits its per-CTA partials in slots [0, gridA); kernel B, later on
// the same stream, deposits at slot_offset = gridA and its last CTA adds all total_slots partials in slot order.
template <int NR, class Op, class Fin>
__device__ __forceinline__ void grid_reduce_ex(double (&v)[NR], double* partials, unsigned int* counter, int slot_offset, int total_slots,
                                               bool finalize, Fin fin) {
    __shared__ double smem[NR * 32];
    __shared__ bool is_last;
    block_reduce<NR, Op>(v, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NR; r++) partials[(size_t)(slot_offset + blockIdx.x) * NR + r] = v[r];
    }
    if (!finalize) return;
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int ticket = atomicInc(counter, gridDim.x - 1);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) acc[r] = Op::ident();
        for (unsigned int b = threadIdx.x; b < (unsigned int)total_slots; b += blockDim.x) {
#pragma unroll
            for (int r = 0; r < NR; r++) acc[r] = Op::comb(acc[r], __ldcg(&partials[(size_t)b * NR + r]));
        }
        block_reduce<NR, Op>(acc, smem);
        if (threadIdx.x == 0) fin(acc);
    }
}
