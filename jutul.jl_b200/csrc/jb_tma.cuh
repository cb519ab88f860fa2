// 1-D bulk asynchronous copies (TMA, cp.async.bulk) global -> shared with mbarrier completion, sm_100a.
//
// Used to stage the contiguous per-chunk streams of the row-chunk kernels (half-face lists, transmissibilities,
// cell records) into shared memory one chunk AHEAD of the arithmetic: one elected thread issues the copies, no
// registers hold data in flight, and the only latency the threads still see is the neighbour gather.
//
// cp.async.bulk needs 16-byte aligned global/shared addresses and a size that is a multiple of 16 B. A chunk's
// element range [i0, i0 + n) of an array of T is therefore widened to the enclosing 16-byte window; the consumer
// indexes the shared tile with the returned lead (elements between the window start and i0). Arrays read this way
// are allocated with >= 16 B of slack (DBuf pads), so the widened window never leaves the allocation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t jb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void jb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(jb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void jb_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void jb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(jb_smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a copy that never lands (a bug) traps instead of hanging the GPU.
__device__ __forceinline__ void jb_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = jb_smem_u32(bar);
    uint32_t ok = 0;
    for (unsigned int spin = 0; spin < (1u << 24); spin++) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ uint64_t jb_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void jb_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(jb_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(jb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void jb_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(jb_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(jb_smem_u32(bar)), "l"(policy) : "memory");
}

// 16-byte window enclosing elements [i0, i0+n) of an array of T
template <class T>
struct JbSpan {
    const char* src;
    uint32_t bytes;
    __device__ __forceinline__ JbSpan(const T* g, size_t i0, int n) {
        const size_t b0 = i0 * sizeof(T);
        const size_t a0 = b0 & ~(size_t)15;
        const size_t a1 = (b0 + (size_t)n * sizeof(T) + 15) & ~(size_t)15;
        src = reinterpret_cast<const char*>(g) + a0;
        bytes = (uint32_t)(a1 - a0);
    }
};
// elements between the window start and element i0
template <class T>
__host__ __device__ __forceinline__ int jb_span_lead(size_t i0) { return (int)(i0 & (16 / sizeof(T) - 1)); }
