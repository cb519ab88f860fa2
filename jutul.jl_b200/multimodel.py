"""MultiModel linear system with reduction = :schur_apply on the device (src/linsolve/multimodel.jl:17-160).

    [B C; D E] [x; y] = [a; b]

B is the device Jacobian of the kept model (reservoir); the eliminated groups (wells, facility) are small scalar sparse
blocks evaluated on the host. The Krylov solver runs on S = B - sum_i C_i E_i^{-1} D_i, preconditioned on B.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

i64 = np.int64
f64 = np.float64


def _pi(a):
    return a.ctypes.data_as(_lib.PI64)


def _pd(a):
    return a.ctypes.data_as(_lib.PF64)


class MultiLinearizedSystemSchur:
    """MultiLinearizedSystem(subsystems; reduction = :schur_apply) (src/multimodel/model.jl:534-601).

    C, D, E: lists (one entry per eliminated group) of (I, J, shape) COO patterns, 1-based, as findnz returns them;
    values are handed over with `update`. C_i is (n bs) x m_i, D_i is m_i x (n bs), E_i is m_i x m_i."""

    def __init__(self, jac, C_patterns, D_patterns, E_patterns):
        self.ctx, self.jac = jac.ctx, jac
        ng = len(E_patterns)
        self.ng = ng
        self.msize = np.array([int(e[2]) for e in E_patterns], dtype=i64)

        def cat(pats):
            ptr = np.ones(ng + 1, dtype=i64)
            for g, pat in enumerate(pats):
                ptr[g + 1] = ptr[g] + len(pat[0])
            I = np.concatenate([np.asarray(pt[0], dtype=i64) for pt in pats]) if ng else np.zeros(0, dtype=i64)
            J = np.concatenate([np.asarray(pt[1], dtype=i64) for pt in pats]) if ng else np.zeros(0, dtype=i64)
            return ptr, np.ascontiguousarray(I), np.ascontiguousarray(J)
        self._C, self._D, self._E = cat(C_patterns), cat(D_patterns), cat(E_patterns)
        h = C.c_void_p()
        check(self.ctx.lib.jb_schur_create(jac.h, ng, _pi(self.msize), _pi(self._C[0]), _pi(self._C[1]), _pi(self._C[2]),
                                           _pi(self._D[0]), _pi(self._D[1]), _pi(self._D[2]), _pi(self._E[0]), _pi(self._E[1]),
                                           _pi(self._E[2]), C.byref(h)), self.ctx.h, "jb_schur_create")
        self.h = h
        self.M = int(self.ctx.lib.jb_schur_size(h))

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.jb_schur_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, C_values, D_values, E_values):
        """get_schur_blocks!(sys, update = true): new block values (lists per group, COO order) and lu!(E_i)."""
        cv = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=f64) for v in C_values]) if self.ng else np.zeros(0))
        dv = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=f64) for v in D_values]) if self.ng else np.zeros(0))
        ev = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=f64) for v in E_values]) if self.ng else np.zeros(0))
        return check(self.ctx.lib.jb_schur_update(self.h, _pd(cv), _pd(dv), _pd(ev)), self.ctx.h, "jb_schur_update")

    def prepare_linear_solve(self, a, b):
        """prepare_linear_solve!: a -= C (E \\ b) in place on the device residual a; b: host vector of the eliminated residuals."""
        b = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=f64) for v in b]) if isinstance(b, (list, tuple)) else b, dtype=f64)
        check(self.ctx.lib.jb_schur_prepare(self.h, a.ptr, _pd(b)), self.ctx.h, "jb_schur_prepare")

    def mul(self, res, x, alpha=1.0, beta=0.0):
        """schur_mul!: res <- beta res + alpha (B x - C (E \\ (D x)))."""
        check(self.ctx.lib.jb_schur_mul(self.h, alpha, x.ptr, beta, res.ptr), self.ctx.h, "jb_schur_mul")
        return res

    def attach(self, krylov):
        """linear_operator(sys): the Krylov solve runs on the Schur complement, the preconditioner on B."""
        check(self.ctx.lib.jb_krylov_set_schur(krylov.h, self.h), self.ctx.h, "jb_krylov_set_schur")

    def detach(self, krylov):
        check(self.ctx.lib.jb_krylov_set_schur(krylov.h, None), self.ctx.h, "jb_krylov_set_schur")

    def update_dx_from_vector(self, dx):
        """schur_dx_update!: dx is the device increment jb_krylov_solve left (-x); returns the eliminated increments y
        split per group (host)."""
        y = np.zeros(self.M)
        check(self.ctx.lib.jb_schur_dx_update(self.h, dx.ptr, _pd(y)), self.ctx.h, "jb_schur_dx_update")
        off = np.concatenate([[0], np.cumsum(self.msize)])
        return [y[off[g]:off[g + 1]].copy() for g in range(self.ng)]
