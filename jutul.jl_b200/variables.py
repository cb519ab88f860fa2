"""Secondary variables and table look-ups on the device (src/variable_evaluation.jl:87-148,260-350, src/interpolation.jl).

The whole dependency graph of a cell is evaluated by one kernel (csrc/variables.cu) with forward-mode partials with
respect to the cell's primary variables; only variables flagged as outputs are written to HBM.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

i64 = np.int64
f64 = np.float64

KINDS = dict(primary=0, parameter=1, const=2, affine=3, product=4, quotient=5, exp=6, power=7, table1d=8, table2d=9)


class VarSpec(C.Structure):
    """jb_var_spec of include/jutul_b200.h"""
    _fields_ = [("kind", C.c_int32), ("dep", C.c_int32 * 3), ("table", C.c_int32), ("output", C.c_int32), ("c", C.c_double * 4)]


def _pd(a):
    return a.ctypes.data_as(_lib.PF64)


class _Table:
    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.jb_table_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        out = (C.c_int64 * 5)()
        check(self.ctx.lib.jb_table_info(self.h, out), self.ctx.h, "jb_table_info")
        return dict(dim=out[0], nx=out[1], ny=out[2], lookup_x=bool(out[3]), lookup_y=bool(out[4]))


def _flag(constant_dx):
    return -1 if constant_dx is None else int(bool(constant_dx))


class LinearInterpolant(_Table):
    """LinearInterpolant(X, F; constant_dx = missing) (src/interpolation.jl:69-99) resident on the device."""

    def __init__(self, ctx, X, F, constant_dx=None):
        self.ctx = ctx
        X = np.ascontiguousarray(X, dtype=f64); F = np.ascontiguousarray(F, dtype=f64)
        if X.shape != F.shape:
            raise ValueError("X and F values must have equal length.")
        h = C.c_void_p()
        check(ctx.lib.jb_table_create_1d(ctx.h, X.shape[0], _pd(X), _pd(F), _flag(constant_dx), C.byref(h)), ctx.h, "jb_table_create_1d")
        self.h = h

    def interpolate(self, x, f, dfdx=None):
        """interpolate(I, x) on device arrays; dfdx optionally receives the slope (the partial ForwardDiff propagates)."""
        check(self.ctx.lib.jb_table_eval(self.h, x.n, x.ptr, None, f.ptr, None if dfdx is None else dfdx.ptr, None), self.ctx.h, "jb_table_eval")
        return f


class BilinearInterpolant(_Table):
    """BilinearInterpolant(xs, ys, fs; constant_dx, constant_dy) (src/interpolation.jl:156-222); fs[i, j] = f(xs[i], ys[j])."""

    def __init__(self, ctx, xs, ys, fs, constant_dx=None, constant_dy=None):
        self.ctx = ctx
        xs = np.ascontiguousarray(xs, dtype=f64); ys = np.ascontiguousarray(ys, dtype=f64)
        fs = np.asarray(fs, dtype=f64)
        if fs.shape != (xs.shape[0], ys.shape[0]):
            raise ValueError("f(x, y) must match lengths of xs (as rows) and ys (as columns)")
        F = np.ascontiguousarray(fs.T).ravel()          # column-major nx x ny
        h = C.c_void_p()
        check(ctx.lib.jb_table_create_2d(ctx.h, xs.shape[0], ys.shape[0], _pd(xs), _pd(ys), _pd(F), _flag(constant_dx), _flag(constant_dy),
                                         C.byref(h)), ctx.h, "jb_table_create_2d")
        self.h = h

    def interpolate(self, x, y, f, dfdx=None, dfdy=None):
        check(self.ctx.lib.jb_table_eval(self.h, x.n, x.ptr, y.ptr, f.ptr, None if dfdx is None else dfdx.ptr,
                                         None if dfdy is None else dfdy.ptr), self.ctx.h, "jb_table_eval")
        return f


def cap_1d(xs, ys, cap_start=True, cap_end=True):
    """End-point capping of get_1d_interpolator (src/interpolation.jl:118-154): constant extrapolation by repeated end values."""
    xs = list(np.asarray(xs, dtype=f64)); ys = list(np.asarray(ys, dtype=f64))
    if len(xs) == 1:
        return np.array([xs[0] - 1.0, xs[0], xs[0] + 1.0]), np.array([ys[0]] * 3)
    if cap_start:
        eps = xs[1] - xs[0]
        xs.insert(0, xs[0] - eps); ys.insert(0, ys[0])
    if cap_end:
        eps = xs[-1] - xs[-2]
        xs.append(xs[-1] + eps); ys.append(ys[-1])
    return np.array(xs), np.array(ys)


def cap_2d(xs, ys, fs, cap_x=(True, True), cap_y=(True, True)):
    """End-point capping of get_2d_interpolator (src/interpolation.jl:224-285)."""
    xs = list(np.asarray(xs, dtype=f64)); ys = list(np.asarray(ys, dtype=f64)); fs = np.asarray(fs, dtype=f64)
    nx, ny = fs.shape
    new = np.zeros((nx + cap_x[0] + cap_x[1], ny + cap_y[0] + cap_y[1]))
    xo, yo = int(cap_x[0]), int(cap_y[0])
    new[xo:nx + xo, yo:ny + yo] = fs
    if cap_x[0]:
        new[0, :] = new[1, :]; xs.insert(0, xs[0] - (xs[1] - xs[0]))
    if cap_x[1]:
        new[-1, :] = new[-2, :]; xs.append(xs[-1] + (xs[-1] - xs[-2]))
    if cap_y[0]:
        new[:, 0] = new[:, 1]; ys.insert(0, ys[0] - (ys[1] - ys[0]))
    if cap_y[1]:
        new[:, -1] = new[:, -2]; ys.append(ys[-1] + (ys[-1] - ys[-2]))
    return np.array(xs), np.array(ys), new


def get_1d_interpolator(ctx, xs, ys, cap_endpoints=True, cap_start=None, cap_end=None, constant_dx=None):
    cs = cap_endpoints if cap_start is None else cap_start
    ce = cap_endpoints if cap_end is None else cap_end
    if len(np.atleast_1d(xs)) == 1 or (cap_endpoints and (cs or ce)):
        xs, ys = cap_1d(xs, ys, cs, ce)
    return LinearInterpolant(ctx, xs, ys, constant_dx=constant_dx)


def get_2d_interpolator(ctx, xs, ys, fs, cap_endpoints=True, constant_dx=None, constant_dy=None):
    if cap_endpoints:
        xs, ys, fs = cap_2d(xs, ys, fs)
    return BilinearInterpolant(ctx, xs, ys, fs, constant_dx=constant_dx, constant_dy=constant_dy)


class SecondaryVariables:
    """The secondary-variable graph of a model, evaluated on the device in sort_secondary_variables! order.

    definitions: ordered dict name -> dict(kind=..., deps=[names], c=[...], table=<interpolant>, output=bool).
    Primaries (kind "primary") carry the partials; parameters (kind "parameter") are value-only inputs."""

    def __init__(self, ctx, nc, definitions):
        self.ctx, self.nc = ctx, int(nc)
        self.names = list(definitions)
        index = {n: i + 1 for i, n in enumerate(self.names)}
        tables = []
        specs = (VarSpec * len(self.names))()
        for i, n in enumerate(self.names):
            d = definitions[n]
            s = specs[i]
            s.kind = KINDS[d["kind"]]
            deps = d.get("deps", [])
            for k in range(3):
                if k < len(deps):
                    if deps[k] not in index:
                        raise KeyError(f"Symbol {deps[k]} must appear exactly once in secondary variables or parameters")
                    s.dep[k] = index[deps[k]]
                else:
                    s.dep[k] = 0
            defaults = dict(affine=[0.0] + [1.0] * len(deps), power=[1.0, 1.0, 0.0, 1.0]).get(d["kind"], [1.0, 0.0, 0.0, 0.0])
            cs = list(d.get("c", []))
            cs = cs + defaults[len(cs):]
            for k in range(4):
                s.c[k] = float(cs[k]) if k < len(cs) else 0.0
            if d.get("table") is not None:
                tables.append(d["table"]); s.table = len(tables)
            else:
                s.table = 0
            s.output = int(bool(d.get("output", False)))
        self._tables = tables
        self.inputs = [n for n in self.names if definitions[n]["kind"] in ("primary", "parameter")]
        self.primaries = [n for n in self.names if definitions[n]["kind"] == "primary"]
        self.outputs = [n for n in self.names if definitions[n].get("output", False)]
        tabs = (C.c_void_p * max(len(tables), 1))(*[t.h for t in tables])
        h = C.c_void_p()
        check(ctx.lib.jb_varprog_create(ctx.h, self.nc, len(self.names), C.cast(specs, C.c_void_p), len(tables),
                                        C.cast(tabs, _lib.PP), C.byref(h)), ctx.h, "jb_varprog_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.jb_varprog_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def order(self):
        """Names in evaluation order (sort_symbols): primaries / parameters included."""
        o = np.zeros(len(self.names), dtype=i64)
        check(self.ctx.lib.jb_varprog_order(self.h, o.ctypes.data_as(_lib.PI64), None), self.ctx.h, "jb_varprog_order")
        return [self.names[k - 1] for k in o]

    def update_secondary_variables(self, state, out):
        """state: name -> DeviceArray (nc) for every primary / parameter; out: name -> DeviceArray ((1 + np) nc) per output."""
        ins = (C.c_void_p * len(self.inputs))(*[state[n].ptr for n in self.inputs])
        outs = (C.c_void_p * max(len(self.outputs), 1))(*[out[n].ptr for n in self.outputs])
        check(self.ctx.lib.jb_varprog_evaluate(self.h, C.cast(ins, _lib.PP), C.cast(outs, _lib.PP)), self.ctx.h, "jb_varprog_evaluate")
        return out
