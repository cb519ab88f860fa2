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
};

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

i64 jb_dist_n_owned(jb_dist* D) { return D->n_owned; }

int jb_dist_halo_launch(jb_dist* D, double* d_vec, int bs) {
    jb_ctx* ctx = D->comm->ctx;
    cudaStream_t st = ctx->stream;
    if (D->nneigh == 0) return JB_OK;
    ProfScope _ps(ctx, JB_PROF_OTHER);
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
int32_t jb_dist_destroy(jb_dist* D) { delete D; return JB_OK; }

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
    int rc = jb_dist_allreduce_launch(D, ctx->d_scalars + 32, n, op);
    if (rc != JB_OK) return rc;
    JB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 64, ctx->d_scalars + 32, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(vals, ctx->h_pinned + 64, n * sizeof(double));
    return JB_OK;
}

}  // extern "C"
