// Argument block of the whole-solve persistent BiCGStab kernel (krylov_persistent.cu) and the layout of the peer-memory
// ("symmetric") buffer of a rank (dist.cu), shared by the two translation units.
#pragma once
#include <stdint.h>

#include "jb_internal.cuh"

// ---- symmetric buffer layout, in 8-byte words (every rank allocates the same header; the vector sections depend on the
//      rank's own sizes, so a sender computes the receiver's offsets from the receiver's n_owned / n_local) ----
#define JB_P2P_MAXW 16
#define JB_P2P_NRED 4
__host__ __device__ inline size_t pk_ar_slot(int parity, int rank) { return ((size_t)parity * JB_P2P_MAXW + rank) * JB_P2P_NRED; }
__host__ __device__ inline size_t pk_ar_flag(int parity, int rank) { return 2 * JB_P2P_MAXW * JB_P2P_NRED + (size_t)parity * JB_P2P_MAXW + rank; }
__host__ __device__ inline size_t pk_halo_flag(int parity, int rank) { return 2 * JB_P2P_MAXW * JB_P2P_NRED + 2 * JB_P2P_MAXW + (size_t)parity * JB_P2P_MAXW + rank; }
// epoch words of the in-kernel halo exchange of the persistent solve: word (kflag_base + source rank)
__host__ __device__ inline size_t pk_kflag_base() { return 2 * JB_P2P_MAXW * JB_P2P_NRED + 4 * JB_P2P_MAXW; }
__host__ __device__ inline size_t pk_stage_base() { return pk_kflag_base() + JB_P2P_MAXW; }
// staging of the host-launched halo exchange: 2 parities x stage_cap x 4 doubles, then the two SpMV operands y and z of the
// Krylov solver ([owned | ghost] cells x up to 4 components each): neighbours store their boundary values straight into
// the ghost sections
inline size_t pk_stage_cap(int64_t n_owned, int64_t n_local) { return (size_t)(n_local - n_owned > 1 ? n_local - n_owned : 1); }
inline size_t pk_ky_base(int64_t n_owned, int64_t n_local) { return pk_stage_base() + 2 * pk_stage_cap(n_owned, n_local) * 4; }
inline size_t pk_kz_base(int64_t n_owned, int64_t n_local) { return pk_ky_base(n_owned, n_local) + (size_t)n_local * 4; }
inline size_t pk_sym_words(int64_t n_owned, int64_t n_local) { return pk_kz_base(n_owned, n_local) + (size_t)n_local * 4; }

// bar: grid-barrier word (bit 31 = phase, low bits = arrivals of the barrier in flight); count: arrivals of a reducing barrier
struct PKSync { unsigned count, bar, push_count, abort; };

struct PKDist {
    int world, rank, nneigh;
    long long nsend;
    double* const* peers;            // every rank's symmetric buffer mapped here (own entry included)
    const int32_t* neigh;            // neighbour ranks
    const int64_t* send_ptr;         // per neighbour: range of send_idx
    const int32_t* send_idx;         // owned local cells to send, grouped by neighbour
    const int64_t* remote_y;         // per neighbour: word offset in ITS symmetric buffer where my first value of y goes (bs = 2)
    const int64_t* remote_z;
    size_t kflag_off;
};

struct PKArgs {
    int n_own, n_id, b0, b1;
    // Jacobian rows [n_id, n_own)
    const S2Chunk* tabA; int nA;
    const int32_t* rowptr; const int32_t* colidx; const double* val;
    // L stream of the B rows, U stream of the R rows
    const S2Chunk* tabL; int cL0, cL1; const int32_t* LptrT; const int32_t* Lcol; const int32_t* forder;
    const S2Chunk* tabU; int cU0, cU1; const int32_t* UptrT; const int32_t* Ucol; const int32_t* border;
    size_t Uoff;
    const double* fv; const double* dinv;
    const int32_t* iso; int n_iso;
    // vectors
    const double* b;
    double *r, *p, *v, *s, *t, *y, *z, *x, *dx;
    double* sc; double* hist; int hist_cap;
    // synchronisation
    PKSync* sync; double* partials;
    unsigned long long* state;       // [0] all-reduce epoch, [1] halo epoch after the solve
    unsigned long long* phase_ns;    // per-phase time accumulators (profiling) or nullptr
    unsigned long long ar_epoch0, halo_epoch0;
    unsigned long long* host_prog;   // pinned host word the kernel reports to: (launch index << 1) | done, or nullptr
    PKDist dist;
};
