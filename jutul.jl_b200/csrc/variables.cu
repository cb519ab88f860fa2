// Secondary-variable evaluation on the device: update_secondary_variables! / update_secondary_variables_state!
// (src/variable_evaluation.jl:87-148) over the dependency graph ordered by sort_secondary_variables! / sort_symbols
// (:260-350), with the table look-ups of src/interpolation.jl (first_lower :3-20, linear_interp :29-40,
// interpolation_constant_lookup :50-67, bilinear_interp :184-216).
//
// The reference walks the sorted OrderedDict and makes one pass over the cells per variable, every array holding ForwardDiff
// Duals (value + one partial per primary variable of the cell). B200 design: the whole DAG is ONE kernel — a thread owns a
// cell, reads its primaries and parameters once, evaluates every secondary variable in dependency order with forward-mode
// partials held in registers / L1-resident local memory (a shared-memory table was measured slower: 1.54 vs 1.13 ms at 10M
// cells, fewer resident warps to hide the pow / exp / division chains), and streams out only the variables flagged as outputs (value plane
// followed by one plane per primary). HBM traffic is inputs + outputs, independent of the depth of the graph; the program
// (a few dozen 64-byte records) and the tables are read through the read-only cache, uniformly across a warp.
//
// The menu is the set of closed forms property models are built from (JutulDarcy's densities, viscosities, relative
// permeabilities, mobilities, masses are compositions of these): affine combinations, products, quotients, exponentials of
// a pressure difference, clamped powers (Brooks-Corey), 1-D and 2-D tables. Anything else stays on the host and enters
// through jb_generic_fill.
#include <algorithm>

#include "jb_internal.cuh"

#define JB_VAR_MAXV 32
#define JB_VAR_MAXP 4

struct jb_table {
    jb_ctx* ctx = nullptr;
    int dim = 1;
    int nx = 0, ny = 0;
    bool lx = false, ly = false;      // constant-spacing lookup available (interpolation_constant_lookup)
    double x0 = 0, dx = 0, y0 = 0, dy = 0;
    std::vector<double> hX, hY, hF;
    DBuf<double> X, Y, F;
};

struct TabDev {   // 64 bytes
    const double* X; const double* Y; const double* F;
    int nx, ny, lx, ly;
    double x0, dx, y0, dy;
};
struct VarOp {    // 64 bytes
    int32_t kind, self, dep[3], table, out_slot, in_slot;
    double c[4];
};

struct jb_varprog {
    jb_ctx* ctx = nullptr;
    i64 nc = 0;
    int nvars = 0, np = 0, nin = 0, nout = 0;
    std::vector<jb_var_spec> specs;
    std::vector<int32_t> order;       // evaluation order over ALL nodes (post-order of the DFS), 0-based
    std::vector<VarOp> ops;           // inputs first (load), then the secondaries in dependency order
    DBuf<VarOp> d_ops;
    DBuf<TabDev> d_tabs;
    DBuf<const double*> d_in;
    DBuf<double*> d_out;
};

// first_lower (interpolation.jl:3-20): constant-spacing lookup, or clamp(searchsortedfirst(tab, x) - 1, 1, n - 1); 0-based result
__device__ __forceinline__ int tab_first_lower(const double* __restrict__ X, int n, bool lookup, double x0, double dx, double x) {
    if (lookup) {
        if (x <= x0 + dx) return 0;
        const int m = (int)floor((x - x0) / dx) + 1;
        return min(m, n - 1) - 1;
    }
    int lo = 0, hi = n;                       // first index with X[i] >= x
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(X + mid) < x) lo = mid + 1; else hi = mid; }
    return max(1, min(lo, n - 1)) - 1;        // (lo + 1) - 1 clamped to [1, n-1], then 0-based
}
// linear_interp_internal: F_0 + dFdX (x - X_0); derivative = dFdX
__device__ __forceinline__ void tab_interp1(const TabDev& T, double x, double& f, double& dfdx) {
    const int ix = tab_first_lower(T.X, T.nx, T.lx != 0, T.x0, T.dx, x);
    const double X0 = __ldg(T.X + ix), F0 = __ldg(T.F + ix);
    dfdx = (__ldg(T.F + ix + 1) - F0) / (__ldg(T.X + ix + 1) - X0);
    f = F0 + dfdx * (x - X0);
}
// bilinear_interp (interpolation.jl:184-216), F column-major nx x ny
__device__ __forceinline__ void tab_interp2(const TabDev& T, double x, double y, double& f, double& dfdx, double& dfdy) {
    const int ix = tab_first_lower(T.X, T.nx, T.lx != 0, T.x0, T.dx, x);
    const int iy = tab_first_lower(T.Y, T.ny, T.ly != 0, T.y0, T.dy, y);
    const double x1 = __ldg(T.X + ix), x2 = __ldg(T.X + ix + 1), y1 = __ldg(T.Y + iy), y2 = __ldg(T.Y + iy + 1);
    const double Dy = y2 - y1;
    const double F11 = __ldg(T.F + ix + (size_t)T.nx * iy), F12 = __ldg(T.F + ix + (size_t)T.nx * (iy + 1));
    const double F21 = __ldg(T.F + ix + 1 + (size_t)T.nx * iy), F22 = __ldg(T.F + ix + 1 + (size_t)T.nx * (iy + 1));
    const double su = (F22 - F12) / (x2 - x1), sl = (F21 - F11) / (x2 - x1);
    const double Fu = F12 + su * (x - x1), Fl = F11 + sl * (x - x1);
    const double wl = (y2 - y) / Dy, wu = (y - y1) / Dy;
    f = wl * Fl + wu * Fu;
    dfdx = wl * sl + wu * su;
    dfdy = (Fu - Fl) / Dy;
}

template <int NP>
__global__ void __launch_bounds__(128) varprog_kernel(i64 nc, int nops, const VarOp* __restrict__ ops, const TabDev* __restrict__ tabs,
                                                      const double* const* __restrict__ in, double* const* __restrict__ out) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) {
        double v[JB_VAR_MAXV];
        double d[JB_VAR_MAXV][NP];
        for (int o = 0; o < nops; o++) {
            const VarOp op = ops[o];
            const int s = op.self;
            double val = 0.0, g[3] = {0.0, 0.0, 0.0};   // value and d/d(dep k)
            int nd = 0;
            switch (op.kind) {
                case JB_VAR_PRIMARY:
                case JB_VAR_PARAMETER:
                    val = __ldg(in[op.in_slot] + c);
                    break;
                case JB_VAR_CONST: val = op.c[0]; break;
                case JB_VAR_AFFINE:
                    val = op.c[0];
                    for (int k = 0; k < 3; k++) if (op.dep[k] >= 0) { val += op.c[1 + k] * v[op.dep[k]]; g[k] = op.c[1 + k]; nd = k + 1; }
                    break;
                case JB_VAR_PRODUCT: {
                    val = op.c[0];
                    for (int k = 0; k < 3; k++) if (op.dep[k] >= 0) { val *= v[op.dep[k]]; nd = k + 1; }
                    for (int k = 0; k < nd; k++) {
                        double t = op.c[0];
                        for (int j = 0; j < nd; j++) if (j != k) t *= v[op.dep[j]];
                        g[k] = t;
                    }
                    break;
                }
                case JB_VAR_QUOTIENT: {
                    const double a = v[op.dep[0]], b = v[op.dep[1]];
                    val = op.c[0] * a / b; g[0] = op.c[0] / b; g[1] = -val / b; nd = 2;
                    break;
                }
                case JB_VAR_EXP: {
                    val = op.c[0] * exp(op.c[1] * (v[op.dep[0]] - op.c[2])); g[0] = op.c[1] * val; nd = 1;
                    break;
                }
                case JB_VAR_POWER: {
                    const double u = (v[op.dep[0]] - op.c[2]) / op.c[3];
                    if (u > 1.0) { val = op.c[0]; }
                    else if (u < 0.0) { val = (op.c[1] == 0.0) ? op.c[0] : 0.0; }
                    else {      // one pow: d/du u^n = n u^(n-1) = n u^n / u for u > 0 (u = 0: n u^(n-1) evaluated directly)
                        const double un = pow(u, op.c[1]);
                        val = op.c[0] * un;
                        g[0] = (op.c[1] == 0.0) ? 0.0 : op.c[0] * op.c[1] * (u > 0.0 ? un / u : pow(u, op.c[1] - 1.0)) / op.c[3];
                    }
                    nd = 1;
                    break;
                }
                case JB_VAR_TABLE1D: {
                    double f, df;
                    tab_interp1(tabs[op.table], v[op.dep[0]], f, df);
                    val = op.c[0] * f; g[0] = op.c[0] * df; nd = 1;
                    break;
                }
                case JB_VAR_TABLE2D: {
                    double f, dfx, dfy;
                    tab_interp2(tabs[op.table], v[op.dep[0]], v[op.dep[1]], f, dfx, dfy);
                    val = op.c[0] * f; g[0] = op.c[0] * dfx; g[1] = op.c[0] * dfy; nd = 2;
                    break;
                }
                default: break;
            }
            v[s] = val;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                double acc = 0.0;
                if (op.kind == JB_VAR_PRIMARY) acc = (q == (int)op.c[0]) ? 1.0 : 0.0;
                else for (int k = 0; k < nd; k++) acc = fma(g[k], d[op.dep[k]][q], acc);
                d[s][q] = acc;
            }
            if (op.out_slot >= 0) {
                double* dst = out[op.out_slot];
                dst[c] = val;
#pragma unroll
                for (int q = 0; q < NP; q++) dst[(size_t)(q + 1) * nc + c] = d[s][q];
            }
        }
    }
}

__global__ void __launch_bounds__(256) table_eval1_kernel(i64 n, TabDev T, const double* __restrict__ x, double* __restrict__ f, double* __restrict__ dfdx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double a, b;
        tab_interp1(T, __ldg(x + i), a, b);
        f[i] = a;
        if (dfdx) dfdx[i] = b;
    }
}
__global__ void __launch_bounds__(256) table_eval2_kernel(i64 n, TabDev T, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ f,
                                                          double* __restrict__ dfdx, double* __restrict__ dfdy) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double a, b, c2;
        tab_interp2(T, __ldg(x + i), __ldg(y + i), a, b, c2);
        f[i] = a;
        if (dfdx) dfdx[i] = b;
        if (dfdy) dfdy[i] = c2;
    }
}

// interpolation_constant_lookup (interpolation.jl:50-67): constant_dx = -1 means `missing` (detect with isapprox)
static bool constant_lookup(const std::vector<double>& X, int flag, double& x0, double& dx) {
    dx = X[1] - X[0]; x0 = X[0];
    bool constant = true;
    if (flag < 0) {
        for (size_t i = 1; i < X.size(); i++) {
            const double a = dx, b = X[i] - X[i - 1];
            constant = constant && std::fabs(a - b) <= 1.4901161193847656e-08 * std::max(std::fabs(a), std::fabs(b));   // isapprox, rtol = sqrt(eps)
        }
    } else constant = flag != 0;
    return constant;
}
static TabDev tab_dev(const jb_table* t) {
    TabDev T;
    T.X = t->X.p; T.Y = t->Y.p; T.F = t->F.p; T.nx = t->nx; T.ny = t->ny; T.lx = t->lx; T.ly = t->ly;
    T.x0 = t->x0; T.dx = t->dx; T.y0 = t->y0; T.dy = t->dy;
    return T;
}
static int vgrid2(jb_ctx* ctx, i64 n, int block) { return (int)std::max<i64>(1, std::min<i64>((n + block - 1) / block, (i64)ctx->sm_count * 16)); }

extern "C" {

// LinearInterpolant(X, F; constant_dx) (interpolation.jl:69-96): X sorted ascending (sorted here if it is not), n >= 2
int32_t jb_table_create_1d(jb_ctx* ctx, int64_t n, const double* X, const double* F, int32_t constant_dx, jb_table** out) {
    if (!ctx || !out || !X || !F || n < 1) return JB_ERR_ARG;
    jb_table* t = new jb_table();
    t->ctx = ctx; t->dim = 1;
    t->hX.assign(X, X + n); t->hF.assign(F, F + n);
    if (n == 1) { t->hX.push_back(X[0] + 1.0); t->hF.push_back(F[0]); }        // single input: constant extrapolation
    if (!std::is_sorted(t->hX.begin(), t->hX.end())) {
        std::vector<size_t> ix(t->hX.size());
        for (size_t i = 0; i < ix.size(); i++) ix[i] = i;
        std::stable_sort(ix.begin(), ix.end(), [&](size_t a, size_t b) { return t->hX[a] < t->hX[b]; });
        std::vector<double> x2(ix.size()), f2(ix.size());
        for (size_t i = 0; i < ix.size(); i++) { x2[i] = t->hX[ix[i]]; f2[i] = t->hF[ix[i]]; }
        t->hX.swap(x2); t->hF.swap(f2);
    }
    t->nx = (int)t->hX.size(); t->ny = 1;
    t->lx = constant_lookup(t->hX, constant_dx, t->x0, t->dx);
    if (t->X.upload(t->hX, ctx->stream) != cudaSuccess || t->F.upload(t->hF, ctx->stream) != cudaSuccess) { delete t; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_table_create_1d: allocation failed"); }
    *out = t;
    return JB_OK;
}
// BilinearInterpolant(xs, ys, fs; constant_dx, constant_dy) (interpolation.jl:156-176): fs is nx x ny, column-major
int32_t jb_table_create_2d(jb_ctx* ctx, int64_t nx, int64_t ny, const double* X, const double* Y, const double* F, int32_t constant_dx,
                           int32_t constant_dy, jb_table** out) {
    if (!ctx || !out || !X || !Y || !F || nx < 2 || ny < 2) return JB_ERR_ARG;
    if (!std::is_sorted(X, X + nx) || !std::is_sorted(Y, Y + ny)) JB_FAIL(ctx, JB_ERR_ARG, "jb_table_create_2d: xs and ys must be sorted");
    jb_table* t = new jb_table();
    t->ctx = ctx; t->dim = 2; t->nx = (int)nx; t->ny = (int)ny;
    t->hX.assign(X, X + nx); t->hY.assign(Y, Y + ny); t->hF.assign(F, F + nx * ny);
    t->lx = constant_lookup(t->hX, constant_dx, t->x0, t->dx);
    t->ly = constant_lookup(t->hY, constant_dy, t->y0, t->dy);
    if (t->X.upload(t->hX, ctx->stream) != cudaSuccess || t->Y.upload(t->hY, ctx->stream) != cudaSuccess || t->F.upload(t->hF, ctx->stream) != cudaSuccess) {
        delete t; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_table_create_2d: allocation failed");
    }
    *out = t;
    return JB_OK;
}
int32_t jb_table_destroy(jb_table* t) { delete t; return JB_OK; }
// info[0] = dim, [1] = nx, [2] = ny, [3] = lookup_x present, [4] = lookup_y present
int32_t jb_table_info(jb_table* t, int64_t* info) {
    if (!t || !info) return JB_ERR_ARG;
    info[0] = t->dim; info[1] = t->nx; info[2] = t->ny; info[3] = t->lx; info[4] = t->ly;
    return JB_OK;
}
// interpolate(I, x) / interpolate(I, x, y) on device arrays; derivative outputs may be NULL
int32_t jb_table_eval(jb_table* t, int64_t n, const double* d_x, const double* d_y, double* d_f, double* d_dfdx, double* d_dfdy) {
    if (!t || n < 0 || !d_x || !d_f || (t->dim == 2 && !d_y)) return JB_ERR_ARG;
    jb_ctx* ctx = t->ctx;
    if (n == 0) return JB_OK;
    {
        ProfScope _ps(ctx, JB_PROF_STATE);
        if (t->dim == 1) table_eval1_kernel<<<vgrid2(ctx, n, 256), 256, 0, ctx->stream>>>(n, tab_dev(t), d_x, d_f, d_dfdx);
        else table_eval2_kernel<<<vgrid2(ctx, n, 256), 256, 0, ctx->stream>>>(n, tab_dev(t), d_x, d_y, d_f, d_dfdx, d_dfdy);
        JB_CHECK_LAUNCH(ctx);
    }
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}

// specs[nvars] in declaration order (primaries, parameters and secondaries in any order); dep entries are 1-based indices
// into specs (0 = unused). The evaluation order is that of sort_symbols (variable_evaluation.jl:327-350): depth-first
// post-order over the nodes in declaration order, dependencies visited in ascending node order (Graphs.jl
// topological_sort_by_dfs, reversed). A cycle or a dangling dependency is an error, as in the reference.
int32_t jb_varprog_create(jb_ctx* ctx, int64_t nc, int32_t nvars, const jb_var_spec* specs, int32_t ntables, jb_table* const* tables,
                          jb_varprog** out) {
    if (!ctx || !out || !specs || nc < 0 || nvars < 1 || ntables < 0) return JB_ERR_ARG;
    if (nvars > JB_VAR_MAXV) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_varprog_create: at most 32 variables per program");
    jb_varprog* P = new jb_varprog();
    P->ctx = ctx; P->nc = nc; P->nvars = nvars;
    P->specs.assign(specs, specs + nvars);
    // adjacency: node -> sorted dependencies
    std::vector<std::vector<int>> adj(nvars);
    for (int i = 0; i < nvars; i++) {
        const jb_var_spec& s = specs[i];
        const bool input = s.kind == JB_VAR_PRIMARY || s.kind == JB_VAR_PARAMETER;
        int need = 0;
        switch (s.kind) {
            case JB_VAR_PRIMARY: case JB_VAR_PARAMETER: case JB_VAR_CONST: need = 0; break;
            case JB_VAR_AFFINE: case JB_VAR_PRODUCT: need = -1; break;   // 1..3
            case JB_VAR_QUOTIENT: case JB_VAR_TABLE2D: need = 2; break;
            case JB_VAR_EXP: case JB_VAR_POWER: case JB_VAR_TABLE1D: need = 1; break;
            default: delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: unknown variable kind");
        }
        int nd = 0;
        for (int k = 0; k < 3; k++) {
            const int dpk = s.dep[k];
            if (dpk < 0 || dpk > nvars) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: dependency index out of range"); }
            if (dpk > 0) { if (k != nd) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: dependencies must be packed from the left"); } nd++; if (!input) adj[i].push_back(dpk - 1); }
        }
        if ((need >= 0 && nd != need) || (need < 0 && nd < 1)) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: wrong number of dependencies for variable " + std::to_string(i + 1)); }
        if ((s.kind == JB_VAR_TABLE1D || s.kind == JB_VAR_TABLE2D)) {
            if (s.table < 1 || s.table > ntables || !tables || !tables[s.table - 1] || tables[s.table - 1]->dim != (s.kind == JB_VAR_TABLE1D ? 1 : 2)) {
                delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: bad table reference");
            }
        }
        if (s.kind == JB_VAR_POWER && s.c[3] == 0.0) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: power law needs a non-zero range c[3]"); }
        std::sort(adj[i].begin(), adj[i].end());
        adj[i].erase(std::unique(adj[i].begin(), adj[i].end()), adj[i].end());
    }
    // topological_sort_by_dfs, post-order
    std::vector<unsigned char> colour(nvars, 0);
    for (int v0 = 0; v0 < nvars; v0++) {
        if (colour[v0]) continue;
        std::vector<int> S(1, v0);
        colour[v0] = 1;
        while (!S.empty()) {
            const int u = S.back();
            int w = -1;
            for (int n : adj[u]) {
                if (colour[n] == 1) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: the variable graph contains a cycle"); }
                if (colour[n] == 0) { w = n; break; }
            }
            if (w >= 0) { colour[w] = 1; S.push_back(w); }
            else { colour[u] = 2; P->order.push_back(u); S.pop_back(); }
        }
    }
    // program: inputs in declaration order, then secondaries in sorted order
    std::vector<int> in_slot(nvars, -1), out_slot(nvars, -1);
    for (int i = 0; i < nvars; i++) {
        if (specs[i].kind == JB_VAR_PRIMARY) { if (P->np >= JB_VAR_MAXP) { delete P; JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_varprog_create: at most 4 primary variables"); } }
        if (specs[i].kind == JB_VAR_PRIMARY || specs[i].kind == JB_VAR_PARAMETER) in_slot[i] = P->nin++;
        if (specs[i].kind == JB_VAR_PRIMARY) P->np++;
    }
    for (int i = 0; i < nvars; i++) if (specs[i].output) out_slot[i] = P->nout++;
    if (P->np < 1) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: no primary variable"); }
    auto make_op = [&](int i, int prim_index) {
        VarOp op;
        const jb_var_spec& s = specs[i];
        op.kind = s.kind; op.self = i;
        for (int k = 0; k < 3; k++) op.dep[k] = s.dep[k] - 1;
        op.table = s.table - 1; op.out_slot = out_slot[i]; op.in_slot = in_slot[i];
        for (int k = 0; k < 4; k++) op.c[k] = s.c[k];
        if (s.kind == JB_VAR_PRIMARY) op.c[0] = (double)prim_index;
        return op;
    };
    int pi = 0;
    for (int i = 0; i < nvars; i++) if (in_slot[i] >= 0) { P->ops.push_back(make_op(i, pi)); if (specs[i].kind == JB_VAR_PRIMARY) pi++; }
    for (int u : P->order) if (in_slot[u] < 0) P->ops.push_back(make_op(u, 0));
    std::vector<TabDev> tabs;
    for (int t = 0; t < ntables; t++) { if (!tables[t]) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_varprog_create: NULL table"); } tabs.push_back(tab_dev(tables[t])); }
    if (tabs.empty()) tabs.push_back(TabDev());
    if (P->d_ops.upload(P->ops, ctx->stream) != cudaSuccess || P->d_tabs.upload(tabs, ctx->stream) != cudaSuccess ||
        P->d_in.alloc(std::max(P->nin, 1)) != cudaSuccess || P->d_out.alloc(std::max(P->nout, 1)) != cudaSuccess) {
        delete P; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_varprog_create: allocation failed");
    }
    *out = P;
    return JB_OK;
}
int32_t jb_varprog_destroy(jb_varprog* P) { delete P; return JB_OK; }
// order[nvars]: 1-based node indices in evaluation order (primaries / parameters included, as sort_symbols returns them);
// counts[0..3] = primaries, inputs, outputs, secondaries
int32_t jb_varprog_order(jb_varprog* P, int64_t* order, int64_t* counts) {
    if (!P) return JB_ERR_ARG;
    if (order) for (int i = 0; i < P->nvars; i++) order[i] = P->order[i] + 1;
    if (counts) { counts[0] = P->np; counts[1] = P->nin; counts[2] = P->nout; counts[3] = P->nvars - P->nin; }
    return JB_OK;
}
// d_inputs: host array of device pointers (nc doubles each), one per PRIMARY / PARAMETER in declaration order;
// d_outputs: host array of device pointers ((1 + np) * nc doubles each: value plane, then d/d(primary q) planes), one per
// variable flagged `output`, in declaration order.
int32_t jb_varprog_evaluate(jb_varprog* P, const double* const* d_inputs, double* const* d_outputs) {
    if (!P || !d_inputs || (P->nout > 0 && !d_outputs)) return JB_ERR_ARG;
    jb_ctx* ctx = P->ctx;
    for (int i = 0; i < P->nin; i++) if (!d_inputs[i]) return JB_ERR_ARG;
    for (int i = 0; i < P->nout; i++) if (!d_outputs[i]) return JB_ERR_ARG;
    cudaStream_t st = ctx->stream;
    JB_CUDA(ctx, cudaMemcpyAsync(P->d_in.p, d_inputs, P->nin * sizeof(double*), cudaMemcpyHostToDevice, st));
    if (P->nout > 0) JB_CUDA(ctx, cudaMemcpyAsync(P->d_out.p, d_outputs, P->nout * sizeof(double*), cudaMemcpyHostToDevice, st));
    JB_CUDA(ctx, cudaStreamSynchronize(st));     // the pointer tables are host stack memory of the caller
    if (P->nc == 0) return JB_OK;
    {
        ProfScope _ps(ctx, JB_PROF_STATE);
        const int g = vgrid2(ctx, P->nc, 128);
        const int nops = (int)P->ops.size();
        switch (P->np) {
            case 1: varprog_kernel<1><<<g, 128, 0, st>>>(P->nc, nops, P->d_ops.p, P->d_tabs.p, P->d_in.p, P->d_out.p); break;
            case 2: varprog_kernel<2><<<g, 128, 0, st>>>(P->nc, nops, P->d_ops.p, P->d_tabs.p, P->d_in.p, P->d_out.p); break;
            case 3: varprog_kernel<3><<<g, 128, 0, st>>>(P->nc, nops, P->d_ops.p, P->d_tabs.p, P->d_in.p, P->d_out.p); break;
            default: varprog_kernel<4><<<g, 128, 0, st>>>(P->nc, nops, P->d_ops.p, P->d_tabs.p, P->d_in.p, P->d_out.p); break;
        }
        JB_CHECK_LAUNCH(ctx);
    }
    JB_CUDA(ctx, cudaStreamSynchronize(st));
    return JB_OK;
}

}  // extern "C"
