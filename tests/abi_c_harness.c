/* Plain-C driver of libjutul_b200.so through dlopen/dlsym — no Python, no torch, no headers beyond include/jutul_b200.h.
 * It is the closest stand-in for Julia's `ccall` this image allows: the calls, their order and their arguments are the ones
 * jutul.jl_b200/julia/JutulB200.jl makes for one Newton iteration of the two-phase model
 *   setup_equation_storage  -> jb_mesh_create          build_sparse_matrix / declare_pattern -> jb_csr_create_tpfa
 *   align_to_jacobian!      -> jb_tpfa_create          setup_twophase!                      -> jb_twophase_create
 *   update_equation!        -> jb_h2d (p, S, M0)       apply_forces_to_equation!            -> jb_twophase_set_sources
 *   update_linearized_system_equation! -> jb_twophase_assemble, jb_d2h(r)
 *   convergence_criterion   -> jb_maxabs_rows
 *   linear_solve!           -> jb_ilu0_create, jb_ilu0_update, jb_krylov_create, jb_krylov_solve, jb_d2h(dx)
 *   update_primary_variable! (device form) -> jb_update_scalar, jb_update_fraction_pair, jb_d2h(p, S)
 *
 * usage: abi_c_harness <libjutul_b200.so> symbols            resolve every entry point used (no GPU needed)
 *        abi_c_harness <libjutul_b200.so> run <in> <out>     run on cuda:0; <in>/<out> are raw little-endian files:
 *   in : int64 nc, nf, nsrc; double dt, rtol; int64 N[2*nf]; double Tf[nf], gdz[nf], pv[nc], params[7], p[nc], s[2nc],
 *        M0[2nc]; int64 src_cells[nsrc]; double src_vals[2*nsrc]
 *   out: int64 nnzb, iters, status; double errors[2], r[2nc], nz[4*nnzb], dx[2nc], p_new[nc], s_new[2nc]            */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/jutul_b200.h"

static double NAN_D(void) { return __builtin_nan(""); }   /* NaN = "no limit" (jutul_b200.h) */

#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 3; }
#define CK(call) do { int32_t rc_ = (call); if (rc_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, p_jb_last_error(ctx)); return 4; } } while (0)

static void* rd(FILE* f, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (bytes && fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(5); }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s lib.so symbols | run in out\n", argv[0]); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(jb_version) SYM(jb_ctx_create) SYM(jb_ctx_destroy) SYM(jb_last_error) SYM(jb_malloc) SYM(jb_free) SYM(jb_h2d) SYM(jb_d2h)
    SYM(jb_mesh_create) SYM(jb_mesh_destroy) SYM(jb_csr_create_tpfa) SYM(jb_csr_destroy) SYM(jb_csr_nnz) SYM(jb_csr_values_get)
    SYM(jb_tpfa_create) SYM(jb_tpfa_destroy) SYM(jb_twophase_create) SYM(jb_twophase_destroy) SYM(jb_twophase_set_sources)
    SYM(jb_twophase_assemble) SYM(jb_maxabs_rows) SYM(jb_ilu0_create) SYM(jb_ilu0_update) SYM(jb_ilu0_destroy) SYM(jb_krylov_create)
    SYM(jb_krylov_solve) SYM(jb_krylov_destroy) SYM(jb_update_scalar) SYM(jb_update_fraction_pair) SYM(jb_launch_count)
    if (strcmp(argv[2], "symbols") == 0) { printf("ok version=%d\n", p_jb_version()); return 0; }
    if (strcmp(argv[2], "run") != 0 || argc < 5) return 2;

    FILE* f = fopen(argv[3], "rb");
    if (!f) { perror("input"); return 5; }
    int64_t hdr[3]; double hd[2];
    if (fread(hdr, 8, 3, f) != 3 || fread(hd, 8, 2, f) != 2) return 5;
    const int64_t nc = hdr[0], nf = hdr[1], nsrc = hdr[2];
    const double dt = hd[0], rtol = hd[1];
    int64_t* N = rd(f, 16 * nf);
    double *Tf = rd(f, 8 * nf), *gdz = rd(f, 8 * nf), *pv = rd(f, 8 * nc), *params = rd(f, 56), *p = rd(f, 8 * nc), *s = rd(f, 16 * nc), *M0 = rd(f, 16 * nc);
    int64_t* src_cells = rd(f, 8 * nsrc);
    double* src_vals = rd(f, 16 * nsrc);
    fclose(f);

    jb_ctx* ctx = NULL;
    if (p_jb_ctx_create(0, &ctx) != 0) { fprintf(stderr, "jb_ctx_create: %s\n", p_jb_last_error(NULL)); return 4; }
    jb_mesh* mesh = NULL; jb_csr* jac = NULL; jb_tpfa* tp = NULL; jb_twophase* law = NULL; jb_ilu* ilu = NULL; jb_krylov* ks = NULL;
    CK(p_jb_mesh_create(ctx, nc, nf, N, &mesh));
    CK(p_jb_csr_create_tpfa(mesh, 2, &jac));
    CK(p_jb_tpfa_create(mesh, jac, &tp));
    CK(p_jb_twophase_create(tp, Tf, gdz, pv, params, &law));
    void *d_p, *d_s, *d_M0, *d_r, *d_dx;
    CK(p_jb_malloc(ctx, 8 * nc, &d_p)); CK(p_jb_malloc(ctx, 16 * nc, &d_s)); CK(p_jb_malloc(ctx, 16 * nc, &d_M0));
    CK(p_jb_malloc(ctx, 16 * nc, &d_r)); CK(p_jb_malloc(ctx, 16 * nc, &d_dx));
    /* update_equation! */
    CK(p_jb_h2d(ctx, d_p, p, 8 * nc)); CK(p_jb_h2d(ctx, d_s, s, 16 * nc)); CK(p_jb_h2d(ctx, d_M0, M0, 16 * nc));
    /* apply_forces! */
    CK(p_jb_twophase_set_sources(law, nsrc, src_cells, src_vals));
    /* update_linearized_system_equation! */
    CK(p_jb_twophase_assemble(law, d_p, d_s, d_M0, dt, d_r));
    const int64_t nnzb = p_jb_csr_nnz(jac);
    double* r = malloc(16 * nc); double* nz = malloc(32 * nnzb); double* dx = malloc(16 * nc);
    CK(p_jb_d2h(ctx, r, d_r, 16 * nc));
    CK(p_jb_csr_values_get(jac, nz));
    /* check_convergence */
    double errors[2];
    CK(p_jb_maxabs_rows(ctx, d_r, 2, nc, errors));
    /* linear_solve!: update_preconditioner!, Krylov, dx = -x */
    CK(p_jb_ilu0_create(jac, NULL, &ilu));
    int32_t st = p_jb_ilu0_update(ilu);
    if (st != 0) { fprintf(stderr, "ilu0_update status %d\n", st); return 6; }
    CK(p_jb_krylov_create(jac, ilu, 0, &ks));
    int32_t iters = 0;
    double hist[1002];
    st = p_jb_krylov_solve(ks, d_r, d_dx, rtol, 1e-12, 1000, 1, 0, &iters, hist, 1002);
    if (st < 0) { fprintf(stderr, "krylov_solve -> %d: %s\n", st, p_jb_last_error(ctx)); return 4; }
    CK(p_jb_d2h(ctx, dx, d_dx, 16 * nc));
    /* update_primary_variables!: Pressure (no limits), Saturations (abs_max 0.2, [0, 1]); dx is 2 x nc, (p, Sw) per cell */
    CK(p_jb_update_scalar(ctx, d_p, d_dx, 2, nc, 1.0, NAN_D(), NAN_D(), NAN_D(), NAN_D(), NAN_D()));
    CK(p_jb_update_fraction_pair(ctx, d_s, (const double*)d_dx + 1, 2, nc, 1.0, 0.2, 0.0, 1.0));
    double* p1 = malloc(8 * nc); double* s1 = malloc(16 * nc);
    CK(p_jb_d2h(ctx, p1, d_p, 8 * nc)); CK(p_jb_d2h(ctx, s1, d_s, 16 * nc));

    FILE* o = fopen(argv[4], "wb");
    if (!o) { perror("output"); return 5; }
    int64_t oh[3] = {nnzb, iters, st};
    fwrite(oh, 8, 3, o); fwrite(errors, 8, 2, o); fwrite(r, 8, 2 * nc, o); fwrite(nz, 8, 4 * nnzb, o); fwrite(dx, 8, 2 * nc, o);
    fwrite(p1, 8, nc, o); fwrite(s1, 8, 2 * nc, o);
    fclose(o);
    printf("ok nc=%lld nnzb=%lld iters=%d status=%d launches=%lld\n", (long long)nc, (long long)nnzb, iters, st, (long long)p_jb_launch_count(ctx));
    p_jb_krylov_destroy(ks); p_jb_ilu0_destroy(ilu); p_jb_twophase_destroy(law); p_jb_tpfa_destroy(tp); p_jb_csr_destroy(jac); p_jb_mesh_destroy(mesh);
    p_jb_free(ctx, d_p); p_jb_free(ctx, d_s); p_jb_free(ctx, d_M0); p_jb_free(ctx, d_r); p_jb_free(ctx, d_dx);
    p_jb_ctx_destroy(ctx);
    return 0;
}
