// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Replaces the PartitionedArrays/MPI layer of the reference (ext/JutulPartitionedArraysExt):
//   consistent!(PVector)          linalg.jl:46, krylov.jl:54,75, interface.jl:200   -> jb_dist_halo_exchange
//   dot / norm on PVector         krylov.jl:107-124 (Allreduce sum)                 -> jb_dist_allreduce (sum)
//   reduce(max, errors)           overloads.jl:194-198                               -> jb_dist_allreduce (max)
// Local cells are numbered [owned | ghost]; the ghost section is ordered by owner rank, so the data received
// from one neighbour is one contiguous range and needs no unpack kernel. The send side packs the boundary
// cells of each neighbour with one gather kernel; all sends/receives of an exchange form one NCCL group on
// the context's stream (no host synchronisation inside the Krylov iteration).
#include <nccl.h>

#include "jb_internal.cuh"
#include "jb_persistent.cuh"

struct jb_comm {
    jb_ctx* ctx;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

struct jb_dist {
    jb_comm* comm;
    i64 n_owned, n_local;
    int nneigh;
    std::vector<int> neigh;                 // neighbour ranks
    std::vector<i64> send_ptr, recv_ptr;    // per neighbour: [send_ptr[k], send_ptr[k+1]) in the pack list; ghost-local range
    DBuf<int32_t> d_send_idx;               // owned local cell ids to pack, grouped by neighbour
    DBuf<double> d_send_buf;                // packed values (max block size 4)
    i64 nsend;
    // ---- peer-memory path (NVLink loads/stores on IPC-mapped buffers; no NCCL on the data path) ----
    bool p2p = false;
    double* sym = nullptr;                  // this rank's symmetric buffer (cudaMalloc, exported through IPC)
    std::vector<double*> peer_sym;          // every rank's symmetric buffer mapped into this process (own entry = sym)
    DBuf<double*> d_peer_sym;
    DBuf<int32_t> d_neigh;                  // neighbour ranks
    DBuf<i64> d_send_ptr, d_recv_ptr, d_remote_off, d_remote_cap;   // per neighbour: remote_off = where my data starts in the neighbour's staging, remote_cap = its ghost count
    DBuf<unsigned int> d_ticket;
    DBuf<int32_t> d_err;
    unsigned long long ar_epoch = 0, halo_epoch = 0;
    unsigned long long k_halo_epoch = 0;    // epoch of the halo exchanges executed inside the persistent Krylov kernel
    i64 stage_cap = 0;                      // ghost cells (staging holds 2 parities x stage_cap x 4 doubles)
    DBuf<i64> d_remote_y, d_remote_z;       // per neighbour: word offset of my values in ITS y / z ghost section (bs = 2)
};

// symmetric buffer layout (in doubles / 8-byte words): jb_persistent.cuh
__host__ __device__ inline size_t p2p_ar_slot(int parity, int rank) { return pk_ar_slot(parity, rank); }
__host__ __device__ inline size_t p2p_ar_flag(int parity, int rank) { return pk_ar_flag(parity, rank); }
__host__ __device__ inline size_t p2p_halo_flag(int parity, int rank) { return pk_halo_flag(parity, rank); }
__host__ __device__ inline size_t p2p_stage_base() { return pk_stage_base(); }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// bounded spin: a lost peer must never hang the GPU (returns false on timeout)
__device__ __forceinline__ bool p2p_wait(const unsigned long long* flag, unsigned long long epoch) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        if (clock64() - t0 > 60000000000LL) return false;   // ~30 s: ranks may reach a collective seconds apart (host skew)
        __nanosleep(40);
    }
    return true;
}

#define JB_NCCL(ctx, call)                                                                         \
    do {                                                                                           \
        ncclResult_t _r = (call);                                                                  \
        if (_r != ncclSuccess) {                                                                   \
            std::string _m = std::string(#call) + ": " + ncclGetErrorString(_r);                   \
            if (ctx) (ctx)->err = _m;                                                              \
            jb_set_global_error(_m);                                                               \
            return JB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

template <int BS>
__global__ void __launch_bounds__(256) halo_pack_kernel(i64 n, const int32_t* __restrict__ idx, const double* __restrict__ v, double* __restrict__ buf) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const size_t c = (size_t)__ldg(idx + i);
#pragma unroll
        for (int e = 0; e < BS; e++) buf[(size_t)i * BS + e] = v[c * BS + e];
    }
}


// ---- fused all-reduce + scalar recurrence over peer memory ------------------------------------------------
// One warp: lane q pushes this rank's partial sums into rank q's slot table (NVLink store), publishes the epoch with a
// release store, then waits for rank q's epoch in the local table. Lane 0 adds the slots in rank order, so every
// rank obtains bitwise the same totals, and applies the Krylov scalar recurrence (jb_krylov_scalars.cuh) in place:
// the all-reduce and its consumer are ONE kernel with no NCCL call and no extra launch.
#include "jb_krylov_scalars.cuh"
__global__ void p2p_allreduce_kernel(double* const* __restrict__ peers, int rank, int world, unsigned long long epoch, double* vals, int n,
                                     int op_max, int fin_which, double* sc, double* hist, int hist_cap, int32_t* err) {
    const int lane = threadIdx.x;
    const int parity = (int)(epoch & 1ULL);
    if (lane < world) {
        double* dst = peers[lane];
        for (int k = 0; k < n; k++) dst[p2p_ar_slot(parity, rank) + k] = vals[k];
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(dst) + p2p_ar_flag(parity, rank), epoch);
    }
    bool ok = true;
    if (lane < world) ok = p2p_wait(reinterpret_cast<unsigned long long*>(peers[rank]) + p2p_ar_flag(parity, lane), epoch);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!ok) *err = 1;
        __threadfence_system();
        const double* mine = peers[rank];
        for (int k = 0; k < n; k++) {
            double a = __ldcv(mine + p2p_ar_slot(parity, 0) + k);
            for (int q = 1; q < world; q++) {
                const double b = __ldcv(mine + p2p_ar_slot(parity, q) + k);
                a = op_max ? ((a != a) ? a : ((b != b) ? b : fmax(a, b))) : a + b;
            }
            vals[k] = a;
        }
        if (fin_which == 0) ks_fin_init(sc, hist);
        else if (fin_which > 0 && sc[KS_DONE] == 0.0) {
            if (fin_which == 1) ks_fin_alpha(sc);
            else if (fin_which == 2) ks_fin_omega(sc);
            else if (fin_which == 3) ks_fin_update2(sc, hist, hist_cap);
        }
    }
}

// ---- halo exchange over peer memory -----------------------------------------------------------------------
// push: every thread stores one boundary cell straight into the neighbour's staging area (NVLink), the last CTA
// publishes the epoch to all neighbours. pull: waits for the neighbours' epochs and copies staging -> ghost section.
template <int BS>
__global__ void __launch_bounds__(256) p2p_halo_push_kernel(double* const* __restrict__ peers, int rank, int nneigh, const int32_t* __restrict__ neigh,
                                                            const i64* __restrict__ send_ptr, const i64* __restrict__ remote_off,
                                                            const int32_t* __restrict__ send_idx, const double* __restrict__ vec, i64 nsend,
                                                            unsigned long long epoch, const i64* __restrict__ remote_cap, unsigned int* ticket) {
    const int parity = (int)(epoch & 1ULL);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nsend; i += (i64)gridDim.x * blockDim.x) {
        int k = 0;
        while (k + 1 < nneigh && i >= send_ptr[k + 1]) k++;
        double* dst = peers[neigh[k]] + p2p_stage_base() + ((size_t)parity * remote_cap[k] + remote_off[k] + (i - send_ptr[k])) * 4;
        const size_t c = (size_t)__ldg(send_idx + i);
#pragma unroll
        for (int e = 0; e < BS; e++) dst[e] = vec[c * BS + e];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
    __syncthreads();
    if (last && (int)threadIdx.x < nneigh) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(peers[neigh[threadIdx.x]]) + p2p_halo_flag(parity, rank), epoch);
    }
}
template <int BS>
__global__ void __launch_bounds__(256) p2p_halo_pull_kernel(double* mine, int nneigh, const int32_t* __restrict__ neigh, i64 n_owned, i64 n_ghost,
                                                            i64 stage_cap, double* __restrict__ vec, unsigned long long epoch, int32_t* err) {
    const int parity = (int)(epoch & 1ULL);
    __shared__ bool ok_s;
    if (threadIdx.x < 32) {
        bool ok = true;
        for (int k = threadIdx.x; k < nneigh; k += 32)
            ok = ok && p2p_wait(reinterpret_cast<unsigned long long*>(mine) + p2p_halo_flag(parity, neigh[k]), epoch);
        ok = __all_sync(0xffffffffu, ok);
        if (threadIdx.x == 0) { ok_s = ok; if (!ok) *err = 2; }
    }
    __syncthreads();
    if (!ok_s) return;
    const double* stage = mine + p2p_stage_base() + (size_t)parity * stage_cap * 4;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n_ghost; g += (i64)gridDim.x * blockDim.x) {
#pragma unroll
        for (int e = 0; e < BS; e++) vec[(n_owned + g) * BS + e] = __ldcv(stage + (size_t)g * 4 + e);
    }
}

// which: 1 = push only (starts a new epoch), 2 = pull only (completes the current epoch), 3 = both
static int p2p_halo_launch(jb_dist* D, double* d_vec, int bs, int which = 3) {
    jb_ctx* ctx = D->comm->ctx;
    cudaStream_t st = ctx->stream;
    const unsigned long long epoch = (which & 1) ? ++D->halo_epoch : D->halo_epoch;
    const i64 n_ghost = D->n_local - D->n_owned;
    const int gp = (int)std::max<i64>(1, std::min<i64>((D->nsend + 255) / 256, (i64)ctx->sm_count * 2));
    const int gq = (int)std::max<i64>(1, std::min<i64>((n_ghost + 255) / 256, (i64)ctx->sm_count * 2));
#define JB_P2P_HALO(BS)                                                                                                                          \
    if (which & 1) {                                                                                                                             \
        p2p_halo_push_kernel<BS><<<gp, 256, 0, st>>>(D->d_peer_sym.p, D->comm->rank, D->nneigh, D->d_neigh.p, D->d_send_ptr.p, D->d_remote_off.p,    \
                                                     D->d_send_idx.p, d_vec, D->nsend, epoch, D->d_remote_cap.p, D->d_ticket.p);                  \
        JB_CHECK_LAUNCH(ctx);                                                                                                                    \
    }                                                                                                                                            \
    if (which & 2) {                                                                                                                             \
        p2p_halo_pull_kernel<BS><<<gq, 256, 0, st>>>(D->sym, D->nneigh, D->d_neigh.p, D->n_owned, n_ghost, D->stage_cap, d_vec, epoch, D->d_err.p); \
        JB_CHECK_LAUNCH(ctx);                                                                                                                    \
    }
    switch (bs) {
        case 1: JB_P2P_HALO(1) break;
        case 2: JB_P2P_HALO(2) break;
        case 3: JB_P2P_HALO(3) break;
        case 4: JB_P2P_HALO(4) break;
        default: return JB_ERR_UNSUPPORTED;
    }
#undef JB_P2P_HALO
    return JB_OK;
}

// all-reduce of d_buf[0..n) fused with the Krylov scalar recurrence `fin_which` (-1: none)
int jb_dist_allreduce_fin_launch(jb_dist* D, double* d_buf, int n, int op_max, int fin_which, double* sc, double* hist, int hist_cap) {
    jb_ctx* ctx = D->comm->ctx;
    if (!D->p2p) return JB_ERR_UNSUPPORTED;
    // one slot of the peer table holds JB_P2P_NRED doubles per (parity, rank): a longer vector would run into the next rank's
    // slot and into the epoch words
    if (n < 1 || n > JB_P2P_NRED) JB_FAIL(ctx, JB_ERR_ARG, "p2p all-reduce: at most JB_P2P_NRED values per exchange");
    ProfScope _ps(ctx, JB_PROF_OTHER);
    const unsigned long long epoch = ++D->ar_epoch;
    p2p_allreduce_kernel<<<1, 32, 0, ctx->stream>>>(D->d_peer_sym.p, D->comm->rank, D->comm->world, epoch, d_buf, n, op_max, fin_which, sc, hist,
                                                    hist_cap, D->d_err.p);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}
bool jb_dist_is_p2p(jb_dist* D) { return D && D->p2p; }

int jb_dist_halo_push_launch(jb_dist* D, double* d_vec, int bs) {
    if (!D->p2p) return JB_ERR_UNSUPPORTED;
    if (D->nneigh == 0) return JB_OK;
    ProfScope _ps(D->comm->ctx, JB_PROF_OTHER);
    return p2p_halo_launch(D, d_vec, bs, 1);
}
int jb_dist_halo_pull_launch(jb_dist* D, double* d_vec, int bs) {
    if (!D->p2p) return JB_ERR_UNSUPPORTED;
    if (D->nneigh == 0) return JB_OK;
    ProfScope _ps(D->comm->ctx, JB_PROF_OTHER);
    return p2p_halo_launch(D, d_vec, bs, 2);
}
i64 jb_dist_n_owned(jb_dist* D) { return D->n_owned; }

int jb_dist_halo_launch(jb_dist* D, double* d_vec, int bs) {
    jb_ctx* ctx = D->comm->ctx;
    cudaStream_t st = ctx->stream;
    if (D->nneigh == 0) return JB_OK;
    ProfScope _ps(ctx, JB_PROF_OTHER);
    if (D->p2p) return p2p_halo_launch(D, d_vec, bs);
    if (D->nsend > 0) {
        const int g = (int)std::max<i64>(1, std::min<i64>((D->nsend + 255) / 256, (i64)ctx->sm_count * 4));
        switch (bs) {
            case 1: halo_pack_kernel<1><<<g, 256, 0, st>>>(D->nsend, D->d_send_idx.p, d_vec, D->d_send_buf.p); break;
            case 2: halo_pack_kernel<2><<<g, 256, 0, st>>>(D->nsend, D->d_send_idx.p, d_vec, D->d_send_buf.p); break;
            case 3: halo_pack_kernel<3><<<g, 256, 0, st>>>(D->nsend, D->d_send_idx.p, d_vec, D->d_send_buf.p); break;
            case 4: halo_pack_kernel<4><<<g, 256, 0, st>>>(D->nsend, D->d_send_idx.p, d_vec, D->d_send_buf.p); break;
            default: return JB_ERR_UNSUPPORTED;
        }
        JB_CHECK_LAUNCH(ctx);
    }
    JB_NCCL(ctx, ncclGroupStart());
    for (int k = 0; k < D->nneigh; k++) {
        const i64 ns = D->send_ptr[k + 1] - D->send_ptr[k], nr = D->recv_ptr[k + 1] - D->recv_ptr[k];
        if (ns > 0) JB_NCCL(ctx, ncclSend(D->d_send_buf.p + D->send_ptr[k] * bs, (size_t)ns * bs, ncclDouble, D->neigh[k], D->comm->comm, st));
        if (nr > 0) JB_NCCL(ctx, ncclRecv(d_vec + (D->n_owned + D->recv_ptr[k]) * bs, (size_t)nr * bs, ncclDouble, D->neigh[k], D->comm->comm, st));
    }
    JB_NCCL(ctx, ncclGroupEnd());
    return JB_OK;
}

int jb_dist_allreduce_launch(jb_dist* D, double* d_buf, int n, int op_max) {
    jb_ctx* ctx = D->comm->ctx;
    if (D->p2p) return jb_dist_allreduce_fin_launch(D, d_buf, n, op_max, -1, nullptr, nullptr, 0);
    ProfScope _ps(ctx, JB_PROF_OTHER);
    JB_NCCL(ctx, ncclAllReduce(d_buf, d_buf, (size_t)n, ncclDouble, op_max ? ncclMax : ncclSum, D->comm->comm, ctx->stream));
    return JB_OK;
}

extern "C" {

int32_t jb_nccl_unique_id(char* id128) {
    if (!id128) return JB_ERR_ARG;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) { jb_set_global_error("ncclGetUniqueId failed"); return JB_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id128, &id, 128);
    return JB_OK;
}

int32_t jb_comm_create(jb_ctx* ctx, int32_t rank, int32_t world, const char* id128, jb_comm** out) {
    if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) return JB_ERR_ARG;
    jb_comm* c = new jb_comm();
    c->ctx = ctx; c->rank = rank; c->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    cudaSetDevice(ctx->device);
    ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; JB_FAIL(ctx, JB_ERR_CUDA, std::string("ncclCommInitRank: ") + ncclGetErrorString(r)); }
    *out = c;
    return JB_OK;
}
int32_t jb_comm_destroy(jb_comm* c) {
    if (c && c->comm) ncclCommDestroy(c->comm);
    delete c;
    return JB_OK;
}

// neigh[nneigh]: neighbour ranks; send_ptr[nneigh+1] / send_idx (1-based local owned cells, grouped by neighbour);
// recv_ptr[nneigh+1]: offsets into the ghost section (ghosts ordered by owner rank).
int32_t jb_dist_create(jb_comm* comm, int64_t n_owned, int64_t n_local, int32_t nneigh, const int32_t* neigh, const int64_t* send_ptr,
                       const int64_t* send_idx, const int64_t* recv_ptr, jb_dist** out) {
    if (!comm || !out || n_owned < 0 || n_local < n_owned || nneigh < 0 || (nneigh > 0 && (!neigh || !send_ptr || !recv_ptr))) return JB_ERR_ARG;
    jb_ctx* ctx = comm->ctx;
    jb_dist* D = new jb_dist();
    D->comm = comm; D->n_owned = n_owned; D->n_local = n_local; D->nneigh = nneigh;
    D->neigh.assign(neigh, neigh + nneigh);
    D->send_ptr.assign(nneigh + 1, 0); D->recv_ptr.assign(nneigh + 1, 0);
    for (int k = 0; k <= nneigh; k++) { D->send_ptr[k] = nneigh ? send_ptr[k] : 0; D->recv_ptr[k] = nneigh ? recv_ptr[k] : 0; }
    D->nsend = D->send_ptr[nneigh];
    if (D->recv_ptr[nneigh] != n_local - n_owned) { delete D; JB_FAIL(ctx, JB_ERR_ARG, "jb_dist_create: receive ranges do not cover the ghost section"); }
    std::vector<int32_t> idx(D->nsend);
    for (i64 i = 0; i < D->nsend; i++) {
        if (send_idx[i] < 1 || send_idx[i] > n_owned) { delete D; JB_FAIL(ctx, JB_ERR_ARG, "jb_dist_create: send index is not an owned cell"); }
        idx[i] = (int32_t)(send_idx[i] - 1);
    }
    if (D->d_send_idx.upload(idx, ctx->stream) != cudaSuccess || D->d_send_buf.alloc((size_t)std::max<i64>(D->nsend, 1) * 4) != cudaSuccess) {
        delete D; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_dist_create: allocation failed");
    }
    *out = D;
    return JB_OK;
}
int32_t jb_dist_destroy(jb_dist* D) {
    if (D) {
        for (size_t q = 0; q < D->peer_sym.size(); q++)
            if (D->peer_sym[q] && D->peer_sym[q] != D->sym) cudaIpcCloseMemHandle(D->peer_sym[q]);
        if (D->sym) cudaFree(D->sym);
    }
    delete D;
    return JB_OK;
}

// Peer-memory setup, step 1: allocate this rank's symmetric buffer and export it (64-byte cudaIpcMemHandle).
int32_t jb_dist_p2p_export(jb_dist* D, char* handle64) {
    if (!D || !handle64) return JB_ERR_ARG;
    jb_ctx* ctx = D->comm->ctx;
    if (D->comm->world > JB_P2P_MAXW) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_dist_p2p_export: world size above JB_P2P_MAXW");
    D->stage_cap = (i64)pk_stage_cap(D->n_owned, D->n_local);
    const size_t words = pk_sym_words(D->n_owned, D->n_local);   // header | staging | y | z (jb_persistent.cuh)
    JB_CUDA(ctx, cudaMalloc((void**)&D->sym, words * sizeof(double)));
    JB_CUDA(ctx, cudaMemset(D->sym, 0, words * sizeof(double)));
    cudaIpcMemHandle_t h;
    JB_CUDA(ctx, cudaIpcGetMemHandle(&h, D->sym));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    memcpy(handle64, &h, 64);
    return JB_OK;
}
// Step 2 (after an all-gather of the handles and of every rank's receive offsets by the launcher):
// handles: world x 64 bytes; remote_off[k] / remote_cap[k]: offset of my data and ghost count in neighbour k's staging.
int32_t jb_dist_p2p_open(jb_dist* D, const char* handles, const int64_t* remote_off, const int64_t* remote_cap, const int64_t* remote_nowned,
                         const int64_t* remote_nlocal) {
    if (!D || !handles || !D->sym || (D->nneigh > 0 && (!remote_off || !remote_cap || !remote_nowned || !remote_nlocal))) return JB_ERR_ARG;
    jb_ctx* ctx = D->comm->ctx;
    const int W = D->comm->world;
    D->peer_sym.assign(W, nullptr);
    for (int q = 0; q < W; q++) {
        if (q == D->comm->rank) { D->peer_sym[q] = D->sym; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) JB_FAIL(ctx, JB_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        D->peer_sym[q] = (double*)p;
    }
    std::vector<int32_t> hn(D->neigh.begin(), D->neigh.end());
    std::vector<i64> ro(remote_off, remote_off + D->nneigh), rc(remote_cap, remote_cap + D->nneigh);
    std::vector<unsigned int> z(1, 0);
    std::vector<int32_t> ze(1, 0);
    cudaStream_t s = ctx->stream;
    // where my boundary values go inside neighbour k's y / z vectors (2 components per cell): its ghost section starts at cell
    // n_owned(k), my range at remote_off[k] inside it
    std::vector<i64> ry(D->nneigh), rz(D->nneigh);
    for (int k = 0; k < D->nneigh; k++) {
        if (remote_nowned[k] < 0 || remote_nlocal[k] < remote_nowned[k] || (i64)pk_stage_cap(remote_nowned[k], remote_nlocal[k]) != remote_cap[k])
            JB_FAIL(ctx, JB_ERR_ARG, "jb_dist_p2p_open: inconsistent sizes of a neighbour");
        ry[k] = (i64)pk_ky_base(remote_nowned[k], remote_nlocal[k]) + 2 * (remote_nowned[k] + remote_off[k]);
        rz[k] = (i64)pk_kz_base(remote_nowned[k], remote_nlocal[k]) + 2 * (remote_nowned[k] + remote_off[k]);
    }
    if (D->nneigh > 0 && (D->d_remote_y.upload(ry, s) != cudaSuccess || D->d_remote_z.upload(rz, s) != cudaSuccess))
        JB_FAIL(ctx, JB_ERR_ALLOC, "jb_dist_p2p_open: allocation failed");
    bool ok = D->d_peer_sym.upload(D->peer_sym, s) == cudaSuccess && D->d_neigh.upload(hn, s) == cudaSuccess &&
              D->d_send_ptr.upload(D->send_ptr, s) == cudaSuccess && D->d_recv_ptr.upload(D->recv_ptr, s) == cudaSuccess &&
              D->d_remote_off.upload(ro, s) == cudaSuccess && D->d_remote_cap.upload(rc, s) == cudaSuccess && D->d_ticket.upload(z, s) == cudaSuccess &&
              D->d_err.upload(ze, s) == cudaSuccess;
    if (!ok) JB_FAIL(ctx, JB_ERR_ALLOC, "jb_dist_p2p_open: allocation failed");
    D->p2p = true;
    return JB_OK;
}
}  // extern "C"

// ---- hooks of the persistent Krylov kernel (krylov_persistent.cu) ----
// the SpMV operands y, z of a distributed solve live in the symmetric buffer ([owned | ghost] x 2 components)
int jb_dist_krylov_vectors(jb_dist* D, i64 n_local, double** y, double** z) {
    if (!D || !D->p2p || !D->sym || n_local != D->n_local) return JB_ERR_ARG;
    *y = D->sym + pk_ky_base(D->n_owned, D->n_local);
    *z = D->sym + pk_kz_base(D->n_owned, D->n_local);
    return JB_OK;
}
int jb_dist_pk_fill(jb_dist* D, PKDist* out, unsigned long long* ar_epoch0, unsigned long long* halo_epoch0) {
    if (!D || !D->p2p) return JB_ERR_UNSUPPORTED;
    out->world = D->comm->world; out->rank = D->comm->rank; out->nneigh = D->nneigh; out->nsend = D->nsend;
    out->peers = D->d_peer_sym.p; out->neigh = D->d_neigh.p; out->send_ptr = D->d_send_ptr.p; out->send_idx = D->d_send_idx.p;
    out->remote_y = D->d_remote_y.p; out->remote_z = D->d_remote_z.p;
    out->kflag_off = pk_kflag_base();
    *ar_epoch0 = D->ar_epoch; *halo_epoch0 = D->k_halo_epoch;
    return JB_OK;
}
void jb_dist_pk_done(jb_dist* D, unsigned long long ar_epoch, unsigned long long halo_epoch) {
    D->ar_epoch = ar_epoch; D->k_halo_epoch = halo_epoch;
}

extern "C" {
// Take the NCCL path from now on (setup fallback: some rank could not export / map the peer buffers and all ranks agreed
// to drop the peer-memory path together). The mapped resources are released with the handle.
int32_t jb_dist_p2p_disable(jb_dist* D) {
    if (!D) return JB_ERR_ARG;
    D->p2p = false;
    return JB_OK;
}
// 0 ok; 1 an all-reduce timed out; 2 a halo exchange timed out (a peer is gone) — the caller must abort the run
int32_t jb_dist_p2p_status(jb_dist* D) {
    if (!D || !D->p2p) return 0;
    int32_t e = 0;
    cudaMemcpyAsync(&e, D->d_err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, D->comm->ctx->stream);
    cudaStreamSynchronize(D->comm->ctx->stream);
    return e;
}

int32_t jb_dist_halo_exchange(jb_dist* D, double* d_vec, int32_t bs) {
    if (!D || !d_vec) return JB_ERR_ARG;
    int rc = jb_dist_halo_launch(D, d_vec, bs);
    if (rc != JB_OK) return rc;
    JB_CUDA(D->comm->ctx, cudaStreamSynchronize(D->comm->ctx->stream));
    return JB_OK;
}
// vals (host, in/out): all-reduced over the ranks; op 0 sum, 1 max
int32_t jb_dist_allreduce(jb_dist* D, double* vals, int32_t n, int32_t op) {
    if (!D || !vals || n < 1 || n > 32) return JB_ERR_ARG;
    jb_ctx* ctx = D->comm->ctx;
    memcpy(ctx->h_pinned + 64, vals, n * sizeof(double));
    JB_CUDA(ctx, cudaMemcpyAsync(ctx->d_scalars + 32, ctx->h_pinned + 64, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // the peer-memory exchange carries JB_P2P_NRED values per epoch: longer vectors go in pieces (every rank cuts alike)
    const int piece = D->p2p ? JB_P2P_NRED : n;
    for (int k0 = 0; k0 < n; k0 += piece) {
        int rc = jb_dist_allreduce_launch(D, ctx->d_scalars + 32 + k0, std::min(piece, n - k0), op);
        if (rc != JB_OK) return rc;
    }
    JB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 64, ctx->d_scalars + 32, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(vals, ctx->h_pinned + 64, n * sizeof(double));
    return JB_OK;
}

}  // extern "C"
