// Micro-benchmark of grid barriers for the fused BiCGStab kernel: ns per barrier for several protocols and grid shapes.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

struct Sync { unsigned count; unsigned gen; unsigned pad[30]; unsigned sub[64 * 32]; };

__device__ unsigned g_abort = 0;
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_add_rel(unsigned* p, unsigned v) { unsigned o; asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }
__device__ __forceinline__ unsigned atom_add_acqrel(unsigned* p, unsigned v) { unsigned o; asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }

__device__ __forceinline__ void spin(Sync* S, unsigned gen) {
    unsigned long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    unsigned n = 0;
    while (ld_acq(&S->gen) == gen) {
        if ((++n & 4095u) == 0) {
            unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (g_abort || t1 - t0 > 2000000000ULL) { g_abort = 1; return; }
        }
    }
}
// variant 0: the current protocol of krylov_persistent.cu
__device__ __forceinline__ void bar_v0(Sync* S, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&S->count, 1u) == gridDim.x - 1) { S->count = 0; __threadfence(); st_rel(&S->gen, gen + 1); }
        else spin(S, gen);
        __threadfence();
    }
    __syncthreads();
    gen++;
}
// variant 1: release atomic / acquire load, no extra fences
__device__ __forceinline__ void bar_v1(Sync* S, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atom_add_acqrel(&S->count, 1u) == gridDim.x - 1) { S->count = 0; st_rel(&S->gen, gen + 1); }
        else spin(S, gen);
    }
    __syncthreads();
    gen++;
}
// variant 2: two-level tree, groups of G CTAs
template <int G>
__device__ __forceinline__ void bar_v2(Sync* S, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned grp = blockIdx.x / G, ngrp = (gridDim.x + G - 1) / G;
        const unsigned gsize = min((unsigned)G, gridDim.x - grp * G);
        bool last = false;
        if (atom_add_acqrel(&S->sub[grp * 32], 1u) == gsize - 1) {
            S->sub[grp * 32] = 0;
            if (atom_add_acqrel(&S->count, 1u) == ngrp - 1) { S->count = 0; st_rel(&S->gen, gen + 1); last = true; }
        }
        if (!last) spin(S, gen);
    }
    __syncthreads();
    gen++;
}
// variant 3: cluster barrier first, then one CTA per cluster in the global protocol
__device__ __forceinline__ void bar_v3(Sync* S, unsigned& gen, unsigned nclusters) {
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();                                   // all CTAs of the cluster arrived (release/acquire at cluster scope)
    if (cl.block_rank() == 0 && threadIdx.x == 0) {
        __threadfence();
        if (atom_add_acqrel(&S->count, 1u) == nclusters - 1) { S->count = 0; st_rel(&S->gen, gen + 1); }
        else spin(S, gen);
        __threadfence();
    }
    cl.sync();
    gen++;
}
// variant 4: cooperative groups grid.sync()
template <int V>
__global__ void __launch_bounds__(128, 6) k_bar(Sync* S, int iters, double* buf, size_t nbuf, int store, unsigned nclusters, unsigned long long* t_out) {
    unsigned gen = *reinterpret_cast<volatile unsigned*>(&S->gen);
    cg::grid_group grid = cg::this_grid();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    unsigned long long t0 = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int it = 0; it < iters && !g_abort; it++) {
        if (store) for (size_t i = tid; i < nbuf; i += nth) buf[i] = (double)it;
        if (V == 0) bar_v0(S, gen);
        else if (V == 1) bar_v1(S, gen);
        else if (V == 2) bar_v2<32>(S, gen);
        else if (V == 5) bar_v2<8>(S, gen);
        else if (V == 3) bar_v3(S, gen, nclusters);
        else grid.sync();
    }
    if (tid == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1)); *t_out = t1 - t0; }
}

template <int V>
static double run(int grid, int cluster, int iters, double* buf, size_t nbuf, int store, Sync* S, unsigned long long* d_t) {
    cudaMemset(S, 0, sizeof(Sync));
    unsigned ncl = cluster > 0 ? grid / cluster : 0;
    void* args[] = {&S, &iters, &buf, &nbuf, &store, &ncl, &d_t};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; na++;
    if (cluster > 0) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; na++; }
    cfg.attrs = at; cfg.numAttrs = na;
    if (cluster > 0) {
        int ncl_max = 0;
        cudaError_t eo = cudaOccupancyMaxActiveClusters(&ncl_max, (const void*)k_bar<V>, &cfg);
        if (eo != cudaSuccess || ncl_max * cluster < grid) { printf(" [cluster%d: max active clusters %d < %d]", cluster, ncl_max, grid / cluster); fflush(stdout); cudaGetLastError(); return -1; }
    }
    unsigned zero = 0; cudaMemcpyToSymbol(g_abort, &zero, 4);
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k_bar<V>, args);
    if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  sync failed: %s\n", cudaGetErrorString(e)); exit(1); }
    unsigned ab = 0; cudaMemcpyFromSymbol(&ab, g_abort, 4);
    if (ab) { printf(" [aborted: barrier timed out]"); fflush(stdout); return -2; }
    unsigned long long t; cudaMemcpy(&t, d_t, 8, cudaMemcpyDeviceToHost);
    return (double)t / iters;
}

int main() {
    Sync* S; cudaMalloc(&S, sizeof(Sync));
    unsigned long long* d_t; cudaMalloc(&d_t, 8);
    const size_t nbuf = 4 << 20;   // 32 MB written per "phase" when store = 1
    double* buf; cudaMalloc(&buf, nbuf * 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs %d\n", p.name, p.multiProcessorCount); fflush(stdout);
    const int iters = 2000;
    for (int store = 0; store <= 1; store++) {
        printf("store=%d (each iteration %s)\n", store, store ? "writes 32 MB then barrier" : "barrier only"); fflush(stdout);
        for (int per : {1, 2, 4, 6}) {
            const int grid = p.multiProcessorCount * per;
            run<0>(grid, 0, 50, buf, nbuf, store, S, d_t);
            printf(" grid %4d:", grid); fflush(stdout);
            printf(" v0 current %.0f ns", run<0>(grid, 0, iters, buf, nbuf, store, S, d_t)); fflush(stdout);
            printf(" | v1 acq_rel %.0f", run<1>(grid, 0, iters, buf, nbuf, store, S, d_t)); fflush(stdout);
            printf(" | v2 tree32 %.0f", run<2>(grid, 0, iters, buf, nbuf, store, S, d_t)); fflush(stdout);
            printf(" | v5 tree8 %.0f", run<5>(grid, 0, iters, buf, nbuf, store, S, d_t)); fflush(stdout);
            printf(" | v4 cg grid.sync %.0f", run<4>(grid, 0, iters, buf, nbuf, store, S, d_t)); fflush(stdout);
            for (int cl : {2, 4, 8}) {
                const int g2 = grid / cl * cl;
                printf(" | v3 cluster%d %.0f", cl, run<3>(g2, cl, iters, buf, nbuf, store, S, d_t)); fflush(stdout);
            }
            printf("\n"); fflush(stdout);
        }
    }
    return 0;
}
