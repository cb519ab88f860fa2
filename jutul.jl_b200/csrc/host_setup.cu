// Setup-time host logic of libjutul_b200.so: topology (half-face CSR), Jacobian
// sparsity, cache->Jacobian alignment and the symbolic phase of ILU(0).
// Integer work done once per model; results are uploaded and stay resident in HBM.
// (The per-Newton-iteration hot path is entirely in the kernel translation units.)
#include <algorithm>
#include <numeric>

#include "jb_internal.cuh"
#include "jb_stream.cuh"
#include "jb_stream2.cuh"

static std::string g_err;
void jb_set_global_error(const std::string& s) { g_err = s; }
const std::string& jb_global_error() { return g_err; }

static bool fits_i32(i64 v) { return v >= 0 && v < (i64)2147483647; }

extern "C" {

// ---------------------------------------------------------------- mesh
// Half-face CSR by counting sort: stable placement in ascending face order per
// cell reproduces "push faces in N order, then sort!" of get_cell_faces
// (src/utils.jl:813-841) without per-cell vectors.
int32_t jb_mesh_create(jb_ctx* ctx, int64_t nc, int64_t nf, const int64_t* N, jb_mesh** out) {
    if (!ctx || !out || nc <= 0 || nf < 0 || (nf > 0 && !N)) JB_FAIL(ctx, JB_ERR_ARG, "jb_mesh_create: bad argument");
    if (!fits_i32(nc) || !fits_i32(2 * nf + 1)) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_mesh_create: index overflow (int32)");
    jb_mesh* m = new jb_mesh();
    m->ctx = ctx; m->nc = nc; m->nf = nf; m->nhf = 2 * nf;
    m->h_left.resize(nf); m->h_right.resize(nf);
    std::vector<int32_t> cnt(nc + 1, 0);
    for (i64 f = 0; f < nf; f++) {
        i64 l = N[2 * f], r = N[2 * f + 1];
        if (l < 1 || l > nc || r < 1 || r > nc || l == r) {
            delete m;
            JB_FAIL(ctx, JB_ERR_ARG, "jb_mesh_create: neighbourship entry out of range or self-loop at face " + std::to_string(f + 1));
        }
        m->h_left[f] = (int32_t)(l - 1); m->h_right[f] = (int32_t)(r - 1);
        cnt[l]++; cnt[r]++;
    }
    m->h_hf_pos.assign(nc + 1, 0);
    for (i64 c = 0; c < nc; c++) m->h_hf_pos[c + 1] = m->h_hf_pos[c] + cnt[c + 1];
    m->h_hf_face.resize(m->nhf); m->h_hf_other.resize(m->nhf); m->h_hf_sign.resize(m->nhf);
    std::vector<int32_t> cur(m->h_hf_pos.begin(), m->h_hf_pos.end() - 1);
    for (i64 f = 0; f < nf; f++) {  // ascending face order => per-cell lists come out sorted
        int32_t l = m->h_left[f], r = m->h_right[f];
        int32_t a = cur[l]++; m->h_hf_face[a] = (int32_t)f; m->h_hf_other[a] = r; m->h_hf_sign[a] = 1;
        int32_t b = cur[r]++; m->h_hf_face[b] = (int32_t)f; m->h_hf_other[b] = l; m->h_hf_sign[b] = -1;
    }
    cudaStream_t s = ctx->stream;
    if (m->d_left.upload(m->h_left, s) != cudaSuccess || m->d_right.upload(m->h_right, s) != cudaSuccess ||
        m->d_hf_pos.upload(m->h_hf_pos, s) != cudaSuccess || m->d_hf_face.upload(m->h_hf_face, s) != cudaSuccess ||
        m->d_hf_other.upload(m->h_hf_other, s) != cudaSuccess || m->d_hf_sign.upload(m->h_hf_sign, s) != cudaSuccess) {
        delete m;
        JB_FAIL(ctx, JB_ERR_ALLOC, "jb_mesh_create: device upload failed");
    }
    *out = m;
    return JB_OK;
}
int32_t jb_mesh_destroy(jb_mesh* m) { delete m; return JB_OK; }

int32_t jb_mesh_halfface(jb_mesh* m, int64_t* faces, int64_t* face_pos, int64_t* cells, int64_t* signs) {
    if (!m) return JB_ERR_ARG;
    for (i64 i = 0; i < m->nhf; i++) {
        if (faces) faces[i] = m->h_hf_face[i] + 1;
        if (cells) cells[i] = m->h_hf_other[i] + 1;
        if (signs) signs[i] = m->h_hf_sign[i];
    }
    if (face_pos) for (i64 c = 0; c <= m->nc; c++) face_pos[c] = m->h_hf_pos[c] + 1;
    return JB_OK;
}

// ---------------------------------------------------------------- csr
static int csr_finish(jb_ctx* ctx, jb_csr* A) {
    A->h_diag.assign(A->n, -1);
    for (i64 r = 0; r < A->n; r++)
        for (int32_t k = A->h_rowptr[r]; k < A->h_rowptr[r + 1]; k++)
            if (A->h_colidx[k] == r) { A->h_diag[r] = k; break; }
    cudaStream_t s = ctx->stream;
    A->h_chunks.clear();
    if (jb_cut_chunks(A->h_rowptr, 0, (int32_t)A->n, A->h_chunks)) A->h_chunks.push_back((int32_t)A->n);
    else A->h_chunks.clear();
    if (!A->h_chunks.empty() && A->d_chunks.upload(A->h_chunks, s) != cudaSuccess) return JB_ERR_ALLOC;
    // TMA-staged SpMV (2x2 blocks)
    A->h_s2.clear(); A->s2_ok = false;
    if (A->bs == 2 && A->n > 0 && jb_s2_cut(A->h_rowptr, 0, (int32_t)A->n, A->h_s2, ctx->sm_count * JB_S2_CTAS_PER_SM)) {
        if (A->d_s2.upload(A->h_s2, s) != cudaSuccess) return JB_ERR_ALLOC;
        A->s2_ok = true;
    }
    if (A->d_rowptr.upload(A->h_rowptr, s) != cudaSuccess || A->d_colidx.upload(A->h_colidx, s) != cudaSuccess ||
        A->d_diag.upload(A->h_diag, s) != cudaSuccess || A->d_val.alloc((size_t)A->nnzb * A->bs * A->bs) != cudaSuccess)
        return JB_ERR_ALLOC;
    if (cudaMemsetAsync(A->d_val.p, 0, (size_t)A->nnzb * A->bs * A->bs * sizeof(double), s) != cudaSuccess) return JB_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) return JB_ERR_CUDA;
    return JB_OK;
}

// Canonical CSR from COO triplets = sparse(J, I, V, n, m) transposed storage
// (src/StaticCSR/mat.jl:73-76): per-row column lists sorted ascending, duplicates
// merged. Two-pass counting sort on rows, then sort+unique inside each row.
int32_t jb_csr_create_from_coo(jb_ctx* ctx, const int64_t* I, const int64_t* J, int64_t nnz_in, int64_t n, int32_t bs,
                               jb_csr** out) {
    if (!ctx || !out || n <= 0 || nnz_in < 0 || bs < 1 || bs > 4) JB_FAIL(ctx, JB_ERR_ARG, "jb_csr_create_from_coo: bad argument");
    // kernels index values as int32 (block index * bs*bs, row * bs): the limit is on the scalar entries, not on the blocks
    if (!fits_i32(n * bs) || !fits_i32(nnz_in * bs * bs)) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_csr_create_from_coo: index overflow (int32)");
    std::vector<int32_t> ptr(n + 1, 0);
    for (i64 k = 0; k < nnz_in; k++) {
        if (I[k] < 1 || I[k] > n || J[k] < 1 || J[k] > n) JB_FAIL(ctx, JB_ERR_ARG, "jb_csr_create_from_coo: index out of range");
        ptr[I[k]]++;
    }
    for (i64 r = 0; r < n; r++) ptr[r + 1] += ptr[r];
    std::vector<int32_t> cols(nnz_in), cur(ptr.begin(), ptr.end() - 1);
    for (i64 k = 0; k < nnz_in; k++) cols[cur[I[k] - 1]++] = (int32_t)(J[k] - 1);
    jb_csr* A = new jb_csr();
    A->ctx = ctx; A->n = n; A->bs = bs;
    A->h_rowptr.assign(n + 1, 0);
    A->h_colidx.reserve(nnz_in);
    for (i64 r = 0; r < n; r++) {
        auto b = cols.begin() + ptr[r], e = cols.begin() + ptr[r + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        A->h_colidx.insert(A->h_colidx.end(), b, e);
        A->h_rowptr[r + 1] = (int32_t)A->h_colidx.size();
    }
    A->nnzb = (i64)A->h_colidx.size();
    int rc = csr_finish(ctx, A);
    if (rc != JB_OK) { delete A; JB_FAIL(ctx, rc, "jb_csr_create_from_coo: device allocation failed"); }
    *out = A;
    return JB_OK;
}

// TPFA pattern {(self, other) for all half-faces} U diagonal
// (src/conservation/conservation.jl:486-505), built directly row by row.
int32_t jb_csr_create_tpfa(jb_mesh* m, int32_t bs, jb_csr** out) {
    if (!m || !out || bs < 1 || bs > 4) return JB_ERR_ARG;
    jb_ctx* ctx = m->ctx;
    if (!fits_i32((m->nhf + m->nc) * (i64)bs * bs)) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_csr_create_tpfa: index overflow (int32)");
    jb_csr* A = new jb_csr();
    A->ctx = ctx; A->n = m->nc; A->bs = bs;
    A->h_rowptr.assign(m->nc + 1, 0);
    A->h_colidx.reserve(m->nhf + m->nc);
    std::vector<int32_t> row;
    for (i64 c = 0; c < m->nc; c++) {
        row.clear();
        row.push_back((int32_t)c);
        for (int32_t i = m->h_hf_pos[c]; i < m->h_hf_pos[c + 1]; i++) row.push_back(m->h_hf_other[i]);
        std::sort(row.begin(), row.end());
        row.erase(std::unique(row.begin(), row.end()), row.end());
        A->h_colidx.insert(A->h_colidx.end(), row.begin(), row.end());
        A->h_rowptr[c + 1] = (int32_t)A->h_colidx.size();
    }
    A->nnzb = (i64)A->h_colidx.size();
    int rc = csr_finish(ctx, A);
    if (rc != JB_OK) { delete A; JB_FAIL(ctx, rc, "jb_csr_create_tpfa: device allocation failed"); }
    *out = A;
    return JB_OK;
}
int32_t jb_csr_destroy(jb_csr* A) { delete A; return JB_OK; }
}  // extern "C"

// ---------------------------------------------------------------- adjoint (transposed) system
// The reference solves adjoint problems by assembling J^T in place: with context = adjoint(ctx) the layout carries
// as_adjoint = true (src/core_types/core_types.jl:140-165), the pattern is transposed (src/models.jl:662-664), every cache
// entry (row, col, eq, partial) is aligned to block (col, row) at in-block index N (eq - 1) + partial
// (src/equations.jl:101-108, find_sparse_position(A, row, col, is_adjoint) :152-161), and the unchanged linear_solve! runs on
// it (src/ad/gradients.jl:519-590). Here the forward assembly kernels keep their layout and J^T is produced from the
// resident J by one gather pass: block k of T reads block tmap[k] of A and swaps its in-block indices.
template <int BS>
__global__ void __launch_bounds__(256) csr_transpose_kernel(i64 nnzb, const int32_t* __restrict__ tmap, const double* __restrict__ src,
                                                            double* __restrict__ dst) {
    constexpr int B2 = BS * BS;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < nnzb; k += (i64)gridDim.x * blockDim.x) {
        const double* a = src + (size_t)__ldg(tmap + k) * B2;
        double v[B2];
#pragma unroll
        for (int q = 0; q < B2; q++) v[q] = __ldg(a + q);
#pragma unroll
        for (int i = 0; i < BS; i++)
#pragma unroll
            for (int j = 0; j < BS; j++) dst[(size_t)k * B2 + j * BS + i] = v[i * BS + j];
    }
}

extern "C" {
int32_t jb_csr_create_transpose(jb_csr* A, jb_csr** out) {
    if (!A || !out) return JB_ERR_ARG;
    jb_ctx* ctx = A->ctx;
    jb_csr* T = new jb_csr();
    T->ctx = ctx; T->n = A->n; T->bs = A->bs; T->nnzb = A->nnzb; T->tr_src = A;
    T->h_rowptr.assign(A->n + 1, 0);
    for (i64 k = 0; k < A->nnzb; k++) T->h_rowptr[A->h_colidx[k] + 1]++;
    for (i64 r = 0; r < A->n; r++) T->h_rowptr[r + 1] += T->h_rowptr[r];
    T->h_colidx.assign(A->nnzb, 0);
    std::vector<int32_t> tmap(A->nnzb), cur(T->h_rowptr.begin(), T->h_rowptr.end() - 1);
    for (i64 r = 0; r < A->n; r++)               // rows ascending: the columns of every row of T come out sorted
        for (int32_t k = A->h_rowptr[r]; k < A->h_rowptr[r + 1]; k++) {
            const int32_t slot = cur[A->h_colidx[k]]++;
            T->h_colidx[slot] = (int32_t)r; tmap[slot] = k;
        }
    int rc = csr_finish(ctx, T);
    if (rc == JB_OK && T->d_tmap.upload(tmap, ctx->stream) != cudaSuccess) rc = JB_ERR_ALLOC;
    if (rc != JB_OK) { delete T; JB_FAIL(ctx, rc, "jb_csr_create_transpose: device allocation failed"); }
    *out = T;
    return JB_OK;
}
int32_t jb_csr_transpose_update(jb_csr* T) {
    if (!T || !T->tr_src) return JB_ERR_ARG;
    jb_ctx* ctx = T->ctx;
    const int g = (int)std::max<i64>(1, std::min<i64>((T->nnzb + 255) / 256, (i64)ctx->sm_count * 16));
    {
        ProfScope _ps(ctx, JB_PROF_OTHER);
        switch (T->bs) {
            case 1: csr_transpose_kernel<1><<<g, 256, 0, ctx->stream>>>(T->nnzb, T->d_tmap.p, T->tr_src->d_val.p, T->d_val.p); break;
            case 2: csr_transpose_kernel<2><<<g, 256, 0, ctx->stream>>>(T->nnzb, T->d_tmap.p, T->tr_src->d_val.p, T->d_val.p); break;
            case 3: csr_transpose_kernel<3><<<g, 256, 0, ctx->stream>>>(T->nnzb, T->d_tmap.p, T->tr_src->d_val.p, T->d_val.p); break;
            case 4: csr_transpose_kernel<4><<<g, 256, 0, ctx->stream>>>(T->nnzb, T->d_tmap.p, T->tr_src->d_val.p, T->d_val.p); break;
            default: return JB_ERR_UNSUPPORTED;
        }
        JB_CHECK_LAUNCH(ctx);
    }
    jb_csr_touch(T);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
}  // extern "C"

// Classify the stream chunks of a rank-local matrix ([owned | ghost] numbering): interior = every row and every column of
// the chunk is owned; boundary = some owned row reads a ghost column (or the chunk straddles n_owned). Chunks of ghost rows
// only are dropped (those rows are -I and never needed by the distributed SpMV).
int jb_csr_split_owned(jb_csr* A, i64 n_owned) {
    if (A->h_chunks.empty()) return JB_ERR_UNSUPPORTED;
    std::vector<int32_t> li, lb;
    const int nch = (int)A->h_chunks.size() - 1;
    for (int c = 0; c < nch; c++) {
        const int32_t r0 = A->h_chunks[c], r1 = A->h_chunks[c + 1];
        if (r0 >= n_owned) continue;
        bool interior = r1 <= n_owned;
        for (int32_t k = A->h_rowptr[r0]; interior && k < A->h_rowptr[r1]; k++) if (A->h_colidx[k] >= n_owned) interior = false;
        (interior ? li : lb).push_back(c);
    }
    A->n_chunks_int = (int)li.size(); A->n_chunks_bnd = (int)lb.size();
    if (li.empty()) li.push_back(0);
    if (lb.empty()) lb.push_back(0);
    if (A->d_chunks_int.upload(li, A->ctx->stream) != cudaSuccess || A->d_chunks_bnd.upload(lb, A->ctx->stream) != cudaSuccess) return JB_ERR_ALLOC;
    A->has_split = true;
    return JB_OK;
}

extern "C" {
int64_t jb_csr_nnz(jb_csr* A) { return A ? A->nnzb : -1; }
int64_t jb_csr_nrows(jb_csr* A) { return A ? A->n : -1; }
int32_t jb_csr_get(jb_csr* A, int64_t* rowptr, int64_t* colidx) {
    if (!A) return JB_ERR_ARG;
    if (rowptr) for (i64 i = 0; i <= A->n; i++) rowptr[i] = (i64)A->h_rowptr[i] + 1;
    if (colidx) for (i64 i = 0; i < A->nnzb; i++) colidx[i] = (i64)A->h_colidx[i] + 1;
    return JB_OK;
}
int32_t jb_csr_values_get(jb_csr* A, double* nz) {
    if (!A || !nz) return JB_ERR_ARG;
    JB_CUDA(A->ctx, cudaMemcpyAsync(nz, A->d_val.p, A->d_val.n * sizeof(double), cudaMemcpyDeviceToHost, A->ctx->stream));
    JB_CUDA(A->ctx, cudaStreamSynchronize(A->ctx->stream));
    return JB_OK;
}
int32_t jb_csr_values_set(jb_csr* A, const double* nz) {
    if (!A || !nz) return JB_ERR_ARG;
    jb_csr_touch(A);
    JB_CUDA(A->ctx, cudaMemcpyAsync(A->d_val.p, nz, A->d_val.n * sizeof(double), cudaMemcpyHostToDevice, A->ctx->stream));
    JB_CUDA(A->ctx, cudaStreamSynchronize(A->ctx->stream));
    return JB_OK;
}
double* jb_csr_values_ptr(jb_csr* A) { return A ? A->d_val.p : nullptr; }
int32_t jb_csr_values_modified(jb_csr* A) {
    if (!A) return JB_ERR_ARG;
    jb_csr_touch(A);
    return JB_OK;
}

// ---------------------------------------------------------------- alignment
static inline int32_t find_in_row(const jb_csr* A, int32_t row, int32_t col) {
    const int32_t* b = A->h_colidx.data() + A->h_rowptr[row];
    const int32_t* e = A->h_colidx.data() + A->h_rowptr[row + 1];
    const int32_t* it = std::lower_bound(b, e, col);
    if (it == e || *it != col) return -1;
    return (int32_t)(it - A->h_colidx.data());
}

int32_t jb_tpfa_create(jb_mesh* m, jb_csr* A, jb_tpfa** out) {
    if (!m || !A || !out || A->n != m->nc) return JB_ERR_ARG;
    jb_ctx* ctx = m->ctx;
    jb_tpfa* t = new jb_tpfa();
    t->mesh = m; t->csr = A;
    t->h_diag_pos.resize(m->nc); t->h_hf_pos.resize(m->nhf); t->h_hf_rowpos.resize(m->nhf);
    bool ok = true;
    for (i64 c = 0; c < m->nc; c++) {
        t->h_diag_pos[c] = A->h_diag[c];
        if (A->h_diag[c] < 0) ok = false;
        for (int32_t i = m->h_hf_pos[c]; i < m->h_hf_pos[c + 1]; i++) {
            int32_t o = m->h_hf_other[i];
            t->h_hf_pos[i] = find_in_row(A, o, (int32_t)c);      // (row = other, col = self)
            t->h_hf_rowpos[i] = find_in_row(A, (int32_t)c, o);   // (row = self, col = other)
            if (t->h_hf_pos[i] < 0 || t->h_hf_rowpos[i] < 0) ok = false;
        }
    }
    if (!ok) { delete t; JB_FAIL(ctx, JB_ERR_ARG, "Jacobian alignment failed: entry not allocated in Jacobian matrix"); }
    if (t->d_hf_pos.upload(t->h_hf_pos, ctx->stream) != cudaSuccess || t->d_hf_rowpos.upload(t->h_hf_rowpos, ctx->stream) != cudaSuccess) {
        delete t; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_tpfa_create: device upload failed");
    }
    *out = t;
    return JB_OK;
}
int32_t jb_tpfa_destroy(jb_tpfa* t) { delete t; return JB_OK; }
int32_t jb_tpfa_positions(jb_tpfa* t, int64_t* diag_pos, int64_t* hf_pos) {
    if (!t) return JB_ERR_ARG;
    if (diag_pos) for (i64 c = 0; c < t->mesh->nc; c++) diag_pos[c] = (i64)t->h_diag_pos[c] + 1;
    if (hf_pos) for (i64 i = 0; i < t->mesh->nhf; i++) hf_pos[i] = (i64)t->h_hf_pos[i] + 1;
    return JB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- ILU(0) symbolic
// Split into strict-lower / strict-upper in-block parts (fixed_block,
// src/StaticCSR/ilu0.jl:13-54) and derive the dependency levels of the row-by-row
// IKJ factorisation (ilu0_factor! :108-144) and of the two triangular sweeps
// (:156-187). Rows of one level are mutually independent; processing levels in
// order and each row's L entries in ascending column order performs exactly the
// arithmetic of the sequential loop. The per-(row, L-entry) update lists replace
// the binary-search getindex U[k, j] of the reference (process_partial_row!).
int jb_ilu_symbolic(jb_ilu* F, const int64_t* partition) {
    const jb_csr* A = F->csr;
    const i64 n = A->n;
    std::vector<int32_t> part;
    if (partition) {
        part.resize(n);
        for (i64 i = 0; i < n; i++) {
            if (partition[i] < 1) return JB_ERR_ARG;
            part[i] = (int32_t)(partition[i] - 1);
        }
    }
    auto same = [&](int32_t a, int32_t b) { return part.empty() || part[a] == part[b]; };
    F->h_Dmap.assign(n, -1);
    std::vector<int32_t> nl(n, 0), nu(n, 0), levF(n, 0), levB(n, 0);
    for (i64 r = 0; r < n; r++) {
        int32_t lf = 0;
        for (int32_t k = A->h_rowptr[r]; k < A->h_rowptr[r + 1]; k++) {
            int32_t c = A->h_colidx[k];
            if (c == r) { F->h_Dmap[r] = k; continue; }
            if (!same((int32_t)r, c)) continue;
            if (c < r) { nl[r]++; lf = std::max(lf, levF[c] + 1); } else nu[r]++;
        }
        if (F->h_Dmap[r] < 0) return JB_ERR_ARG;  // "Diagonal must be present in sparsity pattern."
        levF[r] = lf;
    }
    for (i64 r = n - 1; r >= 0; r--) {
        int32_t lb = 0;
        for (int32_t k = A->h_rowptr[r]; k < A->h_rowptr[r + 1]; k++) {
            int32_t c = A->h_colidx[k];
            if (c > r && same((int32_t)r, c)) lb = std::max(lb, levB[c] + 1);
        }
        levB[r] = lb;
    }
    auto order_by_level = [&](const std::vector<int32_t>& lev, std::vector<int32_t>& order, std::vector<int32_t>& ptr) {
        int32_t nlev = 0;
        for (i64 r = 0; r < n; r++) nlev = std::max(nlev, lev[r] + 1);
        ptr.assign(nlev + 1, 0);
        for (i64 r = 0; r < n; r++) ptr[lev[r] + 1]++;
        for (int32_t l = 0; l < nlev; l++) ptr[l + 1] += ptr[l];
        order.resize(n);
        std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
        for (i64 r = 0; r < n; r++) order[cur[lev[r]]++] = (int32_t)r;  // ascending row index inside a level
        return nlev;
    };
    F->nlevF = order_by_level(levF, F->h_forder, F->h_levF_ptr);
    F->nlevB = order_by_level(levB, F->h_border, F->h_levB_ptr);
    // storage: L rows laid out in forward-level order, U rows in backward-level order
    F->h_Lstart.resize(n); F->h_Lend.resize(n); F->h_Ustart.resize(n); F->h_Uend.resize(n);
    i64 off = 0;
    for (i64 t = 0; t < n; t++) { int32_t r = F->h_forder[t]; F->h_Lstart[r] = (int32_t)off; off += nl[r]; F->h_Lend[r] = (int32_t)off; }
    F->nL = off;
    off = 0;
    for (i64 t = 0; t < n; t++) { int32_t r = F->h_border[t]; F->h_Ustart[r] = (int32_t)off; off += nu[r]; F->h_Uend[r] = (int32_t)off; }
    F->nU = off;
    F->h_Lcol.resize(F->nL); F->h_Lmap.resize(F->nL); F->h_Ucol.resize(F->nU); F->h_Umap.resize(F->nU);
    for (i64 r = 0; r < n; r++) {
        int32_t il = F->h_Lstart[r], iu = F->h_Ustart[r];
        for (int32_t k = A->h_rowptr[r]; k < A->h_rowptr[r + 1]; k++) {
            int32_t c = A->h_colidx[k];
            if (c == r || !same((int32_t)r, c)) continue;
            if (c < r) { F->h_Lcol[il] = c; F->h_Lmap[il++] = k; } else { F->h_Ucol[iu] = c; F->h_Umap[iu++] = k; }
        }
    }
    // update lists. Value buffer layout: [L (nL) | D (n) | U (nU)] in blocks.
    const i64 baseD = F->nL, baseU = F->nL + n;
    F->h_upd_ptr.assign(F->nL + 1, 0);
    F->h_upd_tgt.clear(); F->h_upd_src.clear();
    // Emit in L storage order so that a row's lists are contiguous.
    for (i64 t = 0; t < n; t++) {
        int32_t i = F->h_forder[t];
        for (int32_t li = F->h_Lstart[i]; li < F->h_Lend[i]; li++) {
            int32_t k = F->h_Lcol[li];
            // intersect U row k with {L cols of row i after li} U {i} U {U cols of row i}
            int32_t a = li + 1, ae = F->h_Lend[i];
            int32_t b = F->h_Ustart[i], be = F->h_Uend[i];
            for (int32_t uk = F->h_Ustart[k]; uk < F->h_Uend[k]; uk++) {
                int32_t j = F->h_Ucol[uk];
                if (j < i) {
                    while (a < ae && F->h_Lcol[a] < j) a++;
                    if (a < ae && F->h_Lcol[a] == j) { F->h_upd_tgt.push_back(a); F->h_upd_src.push_back((int32_t)(baseU + uk)); }
                } else if (j == i) {
                    F->h_upd_tgt.push_back((int32_t)(baseD + i)); F->h_upd_src.push_back((int32_t)(baseU + uk));
                } else {
                    while (b < be && F->h_Ucol[b] < j) b++;
                    if (b < be && F->h_Ucol[b] == j) { F->h_upd_tgt.push_back((int32_t)(baseU + b)); F->h_upd_src.push_back((int32_t)(baseU + uk)); }
                }
            }
            F->h_upd_ptr[li + 1] = (int32_t)F->h_upd_tgt.size();
        }
    }
    if (!fits_i32((i64)F->h_upd_tgt.size()) || !fits_i32(F->nL + n + F->nU)) return JB_ERR_UNSUPPORTED;
    // stored-row offsets (storage follows forder / border, so these are plain prefix sums) and stream chunks per level
    F->h_LptrT.resize(n + 1); F->h_UptrT.resize(n + 1);
    for (i64 t = 0; t < n; t++) { F->h_LptrT[t] = F->h_Lstart[F->h_forder[t]]; F->h_UptrT[t] = F->h_Ustart[F->h_border[t]]; }
    F->h_LptrT[n] = (int32_t)F->nL; F->h_UptrT[n] = (int32_t)F->nU;
    F->stream_ok = true;
    auto cut = [&](const std::vector<int32_t>& ptrT, const std::vector<int32_t>& lev_ptr, int nlev, std::vector<int32_t>& chunks,
                   std::vector<int32_t>& lev_chunk) {
        chunks.clear(); lev_chunk.assign(nlev + 1, 0);
        for (int l = 0; l < nlev; l++) {
            lev_chunk[l] = (int32_t)chunks.size();
            if (!jb_cut_chunks(ptrT, lev_ptr[l], lev_ptr[l + 1], chunks)) F->stream_ok = false;
        }
        lev_chunk[nlev] = (int32_t)chunks.size();
        chunks.push_back((int32_t)n);
    };
    cut(F->h_LptrT, F->h_levF_ptr, F->nlevF, F->h_chunksF, F->h_levF_chunk);
    cut(F->h_UptrT, F->h_levB_ptr, F->nlevB, F->h_chunksB, F->h_levB_chunk);
    // tables of the TMA-staged sweeps (2x2 blocks), cut per level as well
    F->s2_ok = (A->bs == 2);
    auto cut2 = [&](const std::vector<int32_t>& ptrT, const std::vector<int32_t>& lev_ptr, int nlev, std::vector<S2Chunk>& tab,
                    std::vector<int32_t>& lev_tab) {
        tab.clear(); lev_tab.assign(nlev + 1, 0);
        for (int l = 0; l < nlev; l++) {
            lev_tab[l] = (int32_t)tab.size();
            if (F->s2_ok && !jb_s2_cut(ptrT, lev_ptr[l], lev_ptr[l + 1], tab, F->csr->ctx->sm_count * JB_S2_CTAS_PER_SM)) F->s2_ok = false;
        }
        lev_tab[nlev] = (int32_t)tab.size();
    };
    if (F->s2_ok) cut2(F->h_LptrT, F->h_levF_ptr, F->nlevF, F->h_s2F, F->h_levF_s2);
    if (F->s2_ok) cut2(F->h_UptrT, F->h_levB_ptr, F->nlevB, F->h_s2B, F->h_levB_s2);
    // two-colour structure: second forward level == rows without U entries, first forward level == second backward level
    // (no row has both L and U entries; rows with neither — e.g. the decoupled ghost rows of a distributed run — are "isolated")
    F->two_colour = false;
    F->h_iso.clear();
    if (F->stream_ok && F->nlevF == 2 && F->nlevB == 2) {
        bool ok = true;
        for (i64 r = 0; ok && r < n; r++) {
            const bool hasL = F->h_Lend[r] != F->h_Lstart[r], hasU = F->h_Uend[r] != F->h_Ustart[r];
            if (hasL && hasU) ok = false;
            if (!hasL && !hasU) F->h_iso.push_back((int32_t)r);
        }
        F->two_colour = ok;
        if (!ok) F->h_iso.clear();
    }
    // stream-form refactorisation of the two-colour case: every update must be "D_i -= m_ik U_ki" (at most one per L entry)
    F->rb_factor = false;
    F->h_usrc.clear();
    if (F->two_colour && A->bs <= 2) {
        bool ok = true;
        F->h_usrc.assign((size_t)F->nL, -1);
        for (i64 r = 0; ok && r < n; r++)
            for (int32_t li = F->h_Lstart[r]; ok && li < F->h_Lend[r]; li++) {
                const int32_t u0 = F->h_upd_ptr[li], u1 = F->h_upd_ptr[li + 1];
                if (u1 - u0 > 1 || (u1 - u0 == 1 && F->h_upd_tgt[u0] != (int32_t)(baseD + r))) ok = false;
                else if (u1 - u0 == 1) F->h_usrc[li] = F->h_upd_src[u0];
            }
        F->rb_factor = ok;
        if (!ok) F->h_usrc.clear();
    }
    // single-pass form of the same refactorisation (ilu.cu, ilu_factor_rb2_kernel): everything is read from the Jacobian
    F->rb2_ok = false;
    if (F->rb_factor && A->bs == 2) {
        const i64 nb = A->nnzb;
        F->h_fdst.assign((size_t)nb, -1);
        for (i64 r = 0; r < n; r++) F->h_fdst[(size_t)F->h_Dmap[r]] = -2;
        for (i64 li = 0; li < F->nL; li++) F->h_fdst[(size_t)F->h_Lmap[li]] = (int32_t)li;
        for (i64 ui = 0; ui < F->nU; ui++) F->h_fdst[(size_t)F->h_Umap[ui]] = (int32_t)(baseU + ui);
        F->h_rtype.assign((size_t)n, 0);
        i64 bmin = n, bmax = -1, nB = 0;
        for (i64 r = 0; r < n; r++)
            if (F->h_Lend[r] != F->h_Lstart[r]) { F->h_rtype[(size_t)r] = 1; bmin = std::min(bmin, r); bmax = std::max(bmax, r); nB++; }
        // gather indices (A_kk, A_ki) per Jacobian block of the second-colour rows, indexed by block - first block of that row range,
        // so that the kernel loads slot and gather indices independently of each other
        F->rb2_kB0 = nB > 0 ? A->h_rowptr[bmin] : 0;
        const i64 kB1 = nB > 0 ? A->h_rowptr[bmax + 1] : 0;
        F->h_gL.assign((size_t)(kB1 - F->rb2_kB0), make_int2(-1, -1));
        for (i64 li = 0; li < F->nL; li++) {
            const int32_t us = F->h_usrc[li];
            F->h_gL[(size_t)(F->h_Lmap[li] - F->rb2_kB0)] = make_int2(F->h_Dmap[F->h_Lcol[li]], us >= 0 ? F->h_Umap[(size_t)(us - baseU)] : -1);
        }
        // row chunks of the Jacobian (<= 128 rows, <= 1024 blocks), never across the border of the second colour's range
        std::vector<int32_t> cuts;
        bool ok = true;
        auto cut_range = [&](i64 a, i64 b) {
            i64 start = a;
            while (ok && start < b) {
                i64 end = start;
                while (end < b && end - start < 128 && A->h_rowptr[end + 1] - A->h_rowptr[start] <= 1024) end++;   // JB_RB2_ROWS / JB_RB2_CAP (ilu.cu)
                if (end == start) { ok = false; break; }
                cuts.push_back((int32_t)start); cuts.push_back((int32_t)end);
                start = end;
            }
        };
        const bool contiguous = nB > 0 && bmax - bmin + 1 == nB;
        if (contiguous) { cut_range(0, bmin); cut_range(bmin, bmax + 1); cut_range(bmax + 1, n); }
        else cut_range(0, n);
        if (ok) {
            const size_t nch = cuts.size() / 2;
            std::vector<size_t> order(nch);
            std::iota(order.begin(), order.end(), (size_t)0);
            if (contiguous && bmin > 0) {
                // a second-colour chunk right after the first-colour chunk at the same relative position: its gathers
                // (A_kk, A_ki of neighbouring first-colour rows) then find those rows in L2
                auto key = [&](size_t c) {
                    const double r0 = cuts[2 * c];
                    if (r0 < bmin) return std::make_pair(r0 / (double)bmin, 0);
                    if (r0 <= bmax) return std::make_pair((r0 - bmin) / (double)nB, 1);
                    return std::make_pair(2.0, 2);
                };
                std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return key(a) < key(b); });
            }
            F->h_rb2_chunks.clear();
            for (size_t c : order) { F->h_rb2_chunks.push_back(cuts[2 * c]); F->h_rb2_chunks.push_back(cuts[2 * c + 1]); }
            F->n_rb2_chunks = (int)nch;
            F->rb2_ok = true;
        }
        if (!F->rb2_ok) { F->h_fdst.clear(); F->h_gL.clear(); F->h_rtype.clear(); F->h_rb2_chunks.clear(); }
    }
    return JB_OK;
}

int jb_ilu_upload(jb_ilu* F) {
    cudaStream_t s = F->csr->ctx->stream;
    // device copies of the row extents are stored in level order (indexed by the position t in forder / border),
    // so a level kernel reads them contiguously
    std::vector<int32_t> ls(F->n), le(F->n), us(F->n), ue(F->n);
    for (i64 t = 0; t < F->n; t++) {
        ls[t] = F->h_Lstart[F->h_forder[t]]; le[t] = F->h_Lend[F->h_forder[t]];
        us[t] = F->h_Ustart[F->h_border[t]]; ue[t] = F->h_Uend[F->h_border[t]];
    }
    bool ok = F->d_forder.upload(F->h_forder, s) == cudaSuccess && F->d_border.upload(F->h_border, s) == cudaSuccess &&
              F->d_Lstart.upload(ls, s) == cudaSuccess && F->d_Lend.upload(le, s) == cudaSuccess &&
              F->d_Ustart.upload(us, s) == cudaSuccess && F->d_Uend.upload(ue, s) == cudaSuccess &&
              F->d_Lcol.upload(F->h_Lcol, s) == cudaSuccess && F->d_Ucol.upload(F->h_Ucol, s) == cudaSuccess &&
              F->d_Lmap.upload(F->h_Lmap, s) == cudaSuccess && F->d_Umap.upload(F->h_Umap, s) == cudaSuccess &&
              F->d_Dmap.upload(F->h_Dmap, s) == cudaSuccess && F->d_upd_ptr.upload(F->h_upd_ptr, s) == cudaSuccess &&
              F->d_upd_tgt.upload(F->h_upd_tgt, s) == cudaSuccess && F->d_upd_src.upload(F->h_upd_src, s) == cudaSuccess;
    ok = ok && (F->h_iso.empty() || F->d_iso.upload(F->h_iso, s) == cudaSuccess);
    if (F->s2_ok) ok = ok && (F->h_s2F.empty() || F->d_s2F.upload(F->h_s2F, s) == cudaSuccess) && (F->h_s2B.empty() || F->d_s2B.upload(F->h_s2B, s) == cudaSuccess);
    ok = ok && (!F->rb_factor || F->h_usrc.empty() || F->d_usrc.upload(F->h_usrc, s) == cudaSuccess);
    if (F->rb2_ok)
        ok = ok && F->d_fdst.upload(F->h_fdst, s) == cudaSuccess && (F->h_gL.empty() || F->d_gL.upload(F->h_gL, s) == cudaSuccess) &&
             F->d_rtype.upload(F->h_rtype, s) == cudaSuccess && F->d_rb2_chunks.upload(F->h_rb2_chunks, s) == cudaSuccess;
    ok = ok && F->d_LptrT.upload(F->h_LptrT, s) == cudaSuccess && F->d_UptrT.upload(F->h_UptrT, s) == cudaSuccess &&
         F->d_chunksF.upload(F->h_chunksF, s) == cudaSuccess && F->d_chunksB.upload(F->h_chunksB, s) == cudaSuccess;
    const size_t b2 = (size_t)F->bs * F->bs;
    ok = ok && F->d_fv.alloc((size_t)(F->nL + F->n + F->nU) * b2) == cudaSuccess && F->d_dinv.alloc((size_t)F->n * b2) == cudaSuccess &&
         F->d_status.alloc(1) == cudaSuccess;
    return ok ? JB_OK : JB_ERR_ALLOC;
}
