// Internal declarations shared by the translation units of libjutul_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jutul_b200.h"

typedef int64_t i64;

#define JB_SM_COUNT_DEFAULT 148

void jb_set_global_error(const std::string& s);

struct jb_ctx {
    int device = 0;
    int sm_count = JB_SM_COUNT_DEFAULT;
    cudaStream_t stream = nullptr;
    std::string err;
    i64 launches = 0;
    // reduction workspace: per-CTA partials + completion counters
    double* d_partials = nullptr;   // JB_MAX_PARTIALS * JB_MAX_RED doubles
    unsigned int* d_counters = nullptr;
    double* d_scalars = nullptr;    // small device scalar scratch (64 doubles)
    double* h_pinned = nullptr;     // pinned host scratch (>= 4096 doubles)
    // optional per-kernel-class timing with CUDA events on `stream` (bench.py roofline pass)
    bool prof_on = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    // per-class times measured inside a kernel (phase timers of the persistent BiCGStab), added by jb_prof_collect
    double prof_extra_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    i64 prof_extra_cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

enum { JB_PROF_STATE = 0, JB_PROF_ASSEMBLY = 1, JB_PROF_SPMV = 2, JB_PROF_ILU_FACTOR = 3, JB_PROF_ILU_APPLY = 4, JB_PROF_VECTOR = 5,
       JB_PROF_NEWTON = 6, JB_PROF_OTHER = 7,
       JB_PROF_FUSED_SOLVE = 8 /* all launches of one fused-kernel BiCGStab solve, event-timed */, JB_PROF_NCLASS = 9 };

void jb_prof_begin(jb_ctx* ctx, int cls);
void jb_prof_end(jb_ctx* ctx);
struct ProfScope {
    jb_ctx* c;
    ProfScope(jb_ctx* ctx, int cls) : c(ctx) { if (c->prof_on) jb_prof_begin(c, cls); }
    ~ProfScope() { if (c->prof_on) jb_prof_end(c); }
};

#define JB_MAX_PARTIALS 2048
#define JB_MAX_RED 4

#define JB_CUDA(ctx, call)                                                                       \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            std::string _m = std::string(#call) + ": " + cudaGetErrorString(_e);                 \
            if (ctx) (ctx)->err = _m;                                                            \
            jb_set_global_error(_m);                                                             \
            return JB_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define JB_FAIL(ctx, code, msg)                  \
    do {                                         \
        std::string _m = (msg);                  \
        if (ctx) (ctx)->err = _m;                \
        jb_set_global_error(_m);                 \
        return (code);                           \
    } while (0)

#define JB_CHECK_LAUNCH(ctx)                                   \
    do {                                                       \
        (ctx)->launches++;                                     \
        JB_CUDA(ctx, cudaGetLastError());                      \
    } while (0)

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        // 16 B of slack: bulk (TMA) copies read the 16-byte window enclosing an element range (jb_tma.cuh)
        return cudaMalloc((void**)&p, count * sizeof(T) + 16);
    }
    cudaError_t upload(const std::vector<T>& h, cudaStream_t s) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        e = cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(s);
    }
};

// chunk table entry of the TMA-staged row-stream kernels (jb_stream2.cuh; 32 B, read as two int4):
// stored rows [t0, t0+nr), blocks [e0, e0+cnt) of the matrix being streamed; flags bit 0 = identity chunk (SpMV)
struct S2Chunk { int32_t t0, nr, e0, cnt, flags, pad0, pad1, pad2; };

struct jb_mesh {
    jb_ctx* ctx;
    i64 nc, nf, nhf;
    std::vector<int32_t> h_left, h_right;                 // per face, 0-based
    std::vector<int32_t> h_hf_pos, h_hf_face, h_hf_other; // half-face CSR, 0-based
    std::vector<int8_t> h_hf_sign;
    DBuf<int32_t> d_left, d_right, d_hf_pos, d_hf_face, d_hf_other;
    DBuf<int8_t> d_hf_sign;
};

struct jb_csr {
    jb_ctx* ctx;
    i64 n, nnzb;
    int bs;
    std::vector<int32_t> h_rowptr, h_colidx, h_diag;  // 0-based
    std::vector<int32_t> h_chunks;                    // row-chunk boundaries for the stream kernels (empty: fallback)
    DBuf<int32_t> d_rowptr, d_colidx, d_diag, d_chunks;
    DBuf<double> d_val;
    // distributed SpMV in two launches (set by jb_csr_split_owned): chunks whose rows and columns are all owned, and chunks
    // with an owned row that reads a ghost column. split_phase: 0 = ordinary single launch, 1 = interior, 2 = boundary.
    DBuf<int32_t> d_chunks_int, d_chunks_bnd;
    int n_chunks_int = 0, n_chunks_bnd = 0;
    int split_phase = 0;
    bool has_split = false;
    // "identity chunks" of the right-preconditioned operator A N^{-1} for a two-colour ILU(0) (krylov.cu): flag per stream chunk,
    // and the vector w the flagged rows copy (set around a launch by the Krylov driver; nullptr = ordinary product)
    // TMA-staged SpMV (bs = 2): chunk table, and a copy with the identity chunks flagged
    std::vector<S2Chunk> h_s2;
    DBuf<S2Chunk> d_s2, d_s2_ident;
    int n_s2_ident = 0;               // entries of d_s2_ident (identity chunks are merged, so it is shorter than h_s2)
    bool s2_ok = false;
    DBuf<unsigned char> d_ident;
    int n_ident_chunks = -1;          // -1 = not analysed yet
    i64 n_ident_rows = 0, n_ident_blocks = 0;
    const void* ident_for = nullptr;  // the jb_ilu the flags were built for
    const double* ident_src = nullptr;
    // generation of the values in d_val: bumped by every entry point that writes the Jacobian (assembly, fill, scaling,
    // values_set, unit_diagonalize). jb_ilu records the generation it was factored from; shortcuts that are exact only for a
    // preconditioner built from the CURRENT values (identity rows of A N^-1, krylov.cu) are taken only when the two match.
    uint64_t val_gen = 1;
    // adjoint system (jb_csr_create_transpose): the matrix this one is the transpose of, and per block the source block
    jb_csr* tr_src = nullptr;
    DBuf<int32_t> d_tmap;
};
inline void jb_csr_touch(jb_csr* A) { if (A) A->val_gen++; }
int jb_csr_split_owned(jb_csr* A, i64 n_owned);
int jb_dist_halo_push_launch(jb_dist* D, double* d_vec, int bs);   // peer-memory path only
int jb_dist_halo_pull_launch(jb_dist* D, double* d_vec, int bs);

struct jb_tpfa {
    jb_mesh* mesh;
    jb_csr* csr;
    std::vector<int32_t> h_diag_pos;    // block index of (c,c)
    std::vector<int32_t> h_hf_pos;      // block index of (other,self)  [reference's table]
    std::vector<int32_t> h_hf_rowpos;   // block index of (self,other)  [row-owner writes]
    DBuf<int32_t> d_hf_pos, d_hf_rowpos;
};

struct jb_perm {
    jb_ctx* ctx;
    i64 n;
    DBuf<int32_t> d_perm;   // device label (0-based) of caller cell i
};

// chunk table entry of the TMA-staged assembly kernel (32 B, read as two int4)
struct Asm2Chunk { int32_t c0, nr, hf0, cnt, row0, flags, pad0, pad1; };

struct jb_twophase {
    jb_tpfa* t;
    jb_perm* perm = nullptr;          // optional caller <-> device renumbering for the host-buffer entry points
    DBuf<double> d_stage;             // staging for permuted host transfers (2 x nc)
    double params[7];
    DBuf<double> d_hf_T, d_hf_sgdz;   // per half-face: T_f and sign*gdz_f
    DBuf<double> d_face_T, d_face_gdz;  // per face (variant B)
    DBuf<int32_t> d_pos_lr, d_pos_rl;   // per face: block index of (l,r) and (r,l) (variant B)
    DBuf<double> d_pv;
    DBuf<double> d_rec;               // per cell {p, sw, rho_w, rho_o}
    DBuf<double> d_rec16;             // per cell and phase {p, rho (v, dp, ds), lambda (v, dp, ds), -}: props_assembly.cu
    DBuf<double> d_src;               // dense 2 x nc source buffer (only if nsrc > 0)
    DBuf<int32_t> d_src_cells;
    DBuf<double> d_src_vals;
    i64 nsrc = 0;
    i64 n_assemble = -1;              // rows to assemble (owned cells of a distributed run); -1 = all
    std::vector<int32_t> h_chunks;    // cell chunks of the stream assembly kernel
    DBuf<int32_t> d_chunks;
    int n_chunks_owned = 0;
    // TMA-staged assembly kernel: chunk table (processing order), 16-bit Jacobian slot of each half-face relative to
    // the first row of its chunk, host copy of the source cells (chunk flags)
    std::vector<Asm2Chunk> h_asm2;
    DBuf<Asm2Chunk> d_asm2;
    DBuf<uint16_t> d_hf_lp;
    std::vector<int32_t> h_src_cells;
    bool asm2_ok = false;
    // resident state for the host-facing perform_step
    DBuf<double> d_p, d_s, d_M0, d_r, d_dx;
    bool M0_resident = false;
};

struct jb_ilu {
    jb_csr* csr;
    i64 n, nL, nU;
    int bs;
    int nlevF, nlevB;
    int diag_kind = 0;                // 0: ILU(0); 1: Jacobi, 2: SPAI(0) (diagprec.cu) — only d_dinv / d_status are used then
    double diag_w = 1.0;
    // host symbolic data (0-based)
    std::vector<int32_t> h_forder, h_border, h_levF_ptr, h_levB_ptr;
    std::vector<int32_t> h_Lstart, h_Lend, h_Ustart, h_Uend;      // per row, offsets into L / U storage
    std::vector<int32_t> h_Lcol, h_Ucol, h_Lmap, h_Umap, h_Dmap;  // storage order
    std::vector<int32_t> h_upd_ptr, h_upd_tgt, h_upd_src;
    std::vector<int32_t> h_LptrT, h_UptrT;            // entry offsets per stored row (level order), n+1 each
    std::vector<int32_t> h_chunksF, h_chunksB;        // stream-kernel chunks (stored-row boundaries), never across a level
    std::vector<int32_t> h_levF_chunk, h_levB_chunk;  // first chunk of each level (nlev+1)
    bool stream_ok = false;
    // TMA-staged sweeps (bs = 2): chunk tables of the L / U storage, first table entry of each level
    std::vector<S2Chunk> h_s2F, h_s2B;
    std::vector<int32_t> h_levF_s2, h_levB_s2;
    DBuf<S2Chunk> d_s2F, d_s2B;
    bool s2_ok = false;
    bool two_colour = false;                          // 2 forward / 2 backward levels with complementary row sets: fused sweeps
    std::vector<int32_t> h_iso;                       // rows with neither L nor U entries (two-colour path)
    DBuf<int32_t> d_iso;
    // two-colour numeric refactorisation in stream form: per L entry the fv index of the one U block that updates the
    // diagonal of its row (-1: none); valid when every update of the factorisation is of that kind
    bool rb_factor = false;
    std::vector<int32_t> h_usrc;
    DBuf<int32_t> d_usrc;
    DBuf<int32_t> d_LptrT, d_UptrT, d_chunksF, d_chunksB;
    // single-pass two-colour refactorisation (ilu_factor_rb2_kernel): per Jacobian block its slot in the factor buffer
    // (-1 dropped coupling, -2 diagonal), per L slot the Jacobian blocks A_kk and A_ki it needs, row type (1 = has L entries),
    // and the row chunks in processing order (first / second colour interleaved by position for L2 reuse)
    bool rb2_ok = false;
    std::vector<int32_t> h_fdst, h_rb2_chunks;
    std::vector<int2> h_gL;
    std::vector<unsigned char> h_rtype;
    DBuf<int32_t> d_fdst, d_rb2_chunks;
    DBuf<int2> d_gL;
    DBuf<unsigned char> d_rtype;
    int n_rb2_chunks = 0;
    int32_t rb2_kB0 = 0;              // first Jacobian block of the second-colour row range (d_gL is indexed relative to it)
    // device
    DBuf<int32_t> d_forder, d_border, d_Lstart, d_Lend, d_Ustart, d_Uend, d_Lcol, d_Ucol, d_Lmap, d_Umap, d_Dmap;
    DBuf<int32_t> d_upd_ptr, d_upd_tgt, d_upd_src;
    DBuf<double> d_fv;     // [L | D | U] blocks
    DBuf<double> d_dinv;   // inverted diagonal blocks
    DBuf<int32_t> d_status;
    uint64_t factored_gen = 0;        // jb_csr::val_gen at the last numeric factorisation (0: never factored)
    cudaGraphExec_t apply_graph = nullptr;
    const double* graph_b = nullptr;
    double* graph_x = nullptr;
};

// ---- multi-GPU: one process per GPU, NCCL over NVLink (dist.cu) ----
struct jb_comm;
struct jb_dist;
int jb_dist_halo_launch(jb_dist* D, double* d_vec, int bs);                 // enqueue on the context's stream
int jb_dist_allreduce_launch(jb_dist* D, double* d_buf, int n, int op_max);  // in place, enqueue only
i64 jb_dist_n_owned(jb_dist* D);
bool jb_dist_is_p2p(jb_dist* D);
int jb_dist_allreduce_fin_launch(jb_dist* D, double* d_buf, int n, int op_max, int fin_which, double* sc, double* hist, int hist_cap);

// host side of the persistent BiCGStab (krylov_persistent.cu): tables built once per (Jacobian, factor, distribution)
struct PKHost {
    bool built = false, ok = false;
    const void* for_ilu = nullptr;
    i64 n_own = 0;
    int n_id = 0, b0 = 0, b1 = 0, nA = 0, n_iso = 0, grid = 0;
    i64 n_id_blocks = 0;
    DBuf<S2Chunk> d_tabA;
    DBuf<int32_t> d_iso;
    DBuf<unsigned> d_sync;
    DBuf<double> d_partials;
    DBuf<unsigned long long> d_state;
    double phase_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // profiling: per-phase device time of the fused kernel
    i64 phase_iters = 0, phase_solves = 0;
};

struct jb_krylov {
    jb_csr* csr;
    jb_ilu* ilu;
    int kind;
    i64 m;  // n*bs
    DBuf<double> r, p, v, s, y, z, t, x, q, c;
    double* yv = nullptr;   // the SpMV operands y = N^-1 p, z = N^-1 s: y.p / z.p, or sections of the rank's peer-memory buffer
    double* zv = nullptr;   // (distributed run: neighbours store their boundary values straight into the ghost sections)
    PKHost pk;
    DBuf<double> d_sc;      // scalar block
    DBuf<double> d_hist;
    double* h_flags = nullptr;  // pinned
    unsigned long long* h_prog = nullptr;    // pinned + mapped: progress word written by the fused kernel (krylov_persistent.cu)
    unsigned long long* d_prog = nullptr;    // its device address
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int hist_cap = 0;
    bool overlap = false;      // interior SpMV overlapped with the halo exchange (JB_OVERLAP=1)
    struct jb_schur* schur = nullptr;   // MultiModel Schur operator S = B - C E^-1 D applied after every SpMV (schur.cu)
    jb_dist* dist = nullptr;   // distributed solve: dots over owned rows + all-reduce, halo exchange before each SpMV
    // gmres: Arnoldi basis (grows on demand), packed Hessenberg/R columns, Givens c/s, z, y
    std::vector<double*> gm_V, gm_Z;  // Z_j = N^{-1} v_j (fgmres only)
    DBuf<double*> gm_Vptr, gm_Zptr;
    int gm_memory = 20;               // Krylov.jl default memory
    bool gm_restart = false;          // Jutul's serial call: restart = false; its distributed call: true
    bool gm_flexible = false;         // fgmres!
    DBuf<double> gm_R, gm_cs;
};

struct jb_schur;
int jb_schur_correct_launch(jb_schur* S, const double* d_x, double* d_res, double alpha);   // res -= alpha C (E \ (D x))
i64 jb_schur_nrows(jb_schur* S);

// ---- launchers implemented in the kernel translation units (all enqueue on ctx->stream) ----
enum { JB_DOT_NONE = 0, JB_DOT_CV = 1, JB_DOT_TS_TT = 2 };

int jb_launch_spmv(jb_csr* A, double alpha, const double* d_x, double beta, double* d_y);
// y = A*x fused with dot products; results accumulated by the finalizer into sc (see krylov.cu)
int jb_launch_spmv_dots(jb_csr* A, const double* d_x, double* d_y, int mode, const double* d_u, double* d_sc, i64 n_dot = -1);

int jb_ilu_symbolic(jb_ilu* F, const int64_t* partition);
int jb_ilu_upload(jb_ilu* F);
int jb_launch_ilu_factor(jb_ilu* F);
int jb_launch_ilu_apply(jb_ilu* F, const double* d_b, double* d_x);
int jb_launch_diag_factor(jb_ilu* F);
int jb_launch_diag_apply(jb_ilu* F, const double* d_b, double* d_x, const double* d_sc);

int jb_launch_twophase_state(jb_twophase* m, const double* d_p, const double* d_s);
int jb_launch_twophase_assemble(jb_twophase* m, const double* d_M0, double dt, double* d_r, bool jac);
int jb_launch_twophase_assemble_faces(jb_twophase* m, const double* d_M0, double dt, double* d_r);
int jb_launch_twophase_mass(jb_twophase* m, const double* d_p, const double* d_s, double* d_M);

// small dense block helpers usable on host and device
#ifdef __CUDACC__
#define JB_HD __host__ __device__ __forceinline__
#else
#define JB_HD inline
#endif

template <int BS>
JB_HD void blk_inv(const double* A, double* B) {
    if (BS == 1) {
        B[0] = 1.0 / A[0];
    } else if (BS == 2) {
        double det = A[0] * A[3] - A[2] * A[1];
        double idet = 1.0 / det;
        B[0] = A[3] * idet; B[1] = -(A[1] * idet); B[2] = -(A[2] * idet); B[3] = A[0] * idet;
    } else {
        double M[BS * BS], I[BS * BS];
        for (int i = 0; i < BS * BS; i++) { M[i] = A[i]; I[i] = 0.0; }
        for (int i = 0; i < BS; i++) I[i * BS + i] = 1.0;
        for (int c = 0; c < BS; c++) {
            int piv = c;
            for (int r = c + 1; r < BS; r++) if (fabs(M[c * BS + r]) > fabs(M[c * BS + piv])) piv = r;
            if (piv != c) for (int k = 0; k < BS; k++) {
                double a = M[k * BS + c]; M[k * BS + c] = M[k * BS + piv]; M[k * BS + piv] = a;
                a = I[k * BS + c]; I[k * BS + c] = I[k * BS + piv]; I[k * BS + piv] = a;
            }
            double ip = 1.0 / M[c * BS + c];
            for (int k = 0; k < BS; k++) { M[k * BS + c] *= ip; I[k * BS + c] *= ip; }
            for (int r = 0; r < BS; r++) if (r != c) {
                double f = M[c * BS + r];
                for (int k = 0; k < BS; k++) { M[k * BS + r] -= f * M[k * BS + c]; I[k * BS + r] -= f * I[k * BS + c]; }
            }
        }
        for (int i = 0; i < BS * BS; i++) B[i] = I[i];
    }
}
// C = A*B, column-major
template <int BS>
JB_HD void blk_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int j = 0; j < BS; j++)
#pragma unroll
        for (int i = 0; i < BS; i++) {
            double s = A[i] * B[j * BS];
#pragma unroll
            for (int p = 1; p < BS; p++) s += A[p * BS + i] * B[j * BS + p];
            C[j * BS + i] = s;
        }
}
// y -= A*x
template <int BS>
JB_HD void blk_submulvec(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < BS; i++) {
        double s = A[i] * x[0];
#pragma unroll
        for (int p = 1; p < BS; p++) s += A[p * BS + i] * x[p];
        y[i] -= s;
    }
}
template <int BS>
JB_HD void blk_mulvec(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < BS; i++) {
        double s = A[i] * x[0];
#pragma unroll
        for (int p = 1; p < BS; p++) s += A[p * BS + i] * x[p];
        y[i] = s;
    }
}
