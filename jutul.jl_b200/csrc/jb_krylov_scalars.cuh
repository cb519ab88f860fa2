// Device-resident scalar block of the Krylov solvers. Kernels read and update it
// in place so that an iteration needs no host round trip.
#pragma once
enum {
    KS_RHO = 0,       // rho_k = <c, r_{k-1}>
    KS_ALPHA = 1,
    KS_OMEGA = 2,
    KS_BETA = 3,
    KS_RNORM = 4,
    KS_EPS = 5,       // atol + rtol*|r0|
    KS_R0 = 6,
    KS_DONE = 7,      // != 0: all later kernels of this solve are no-ops
    KS_STATUS = 8,    // 0 solved, 1 running/itmax, 2 breakdown
    KS_ITER = 9,
    KS_ABS_TOL = 10,  // manual (min_it > 1) criterion
    KS_REL_TOL = 11,
    KS_MIN_IT = 12,
    KS_MANUAL = 13,
    KS_ITMAX = 14,
    KS_RTOL = 15,
    KS_ATOL = 16,
    KS_SIZE = 32
};
