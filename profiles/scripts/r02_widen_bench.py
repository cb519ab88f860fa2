"""Round 2: device times of the kernels either side of the Newton hot path (CUDA events on the library's stream), with the
compulsory bytes each one moves. Usage: python profiles/scripts/r02_widen_bench.py [nx ny nz]  -> JSON lines on stdout."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

J = g.load_package()
dims = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else [216, 216, 216]
PEAK = 6535.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ctx = J.B200Context(0)


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    J.timer_start(ctx)
    for _ in range(reps):
        fn()
    return J.timer_stop(ctx) / reps


def line(name, ms, nbytes, **kw):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel=name, ms=round(ms, 4), compulsory_bytes=int(nbytes), gbs=round(gbs, 1), frac_of_measured_peak=round(gbs / PEAK, 3), **kw)), flush=True)


w = J.workloads.unstructured_hex(*dims)
nc, nf = w["nc"], w["nf"]
rng = np.random.default_rng(1)

# ---- secondary-variable graph: 2 primaries + 1 parameter in, 5 outputs x (1 + 2) planes out
xs = np.linspace(5e4, 4e5, 64)
t1 = J.get_1d_interpolator(ctx, xs, 1e-3 * (1 + 1e-6 * xs))
gx = np.linspace(5e4, 4e5, 32); gy = np.linspace(0.0, 1.0, 16)
t2 = J.get_2d_interpolator(ctx, gx, gy, 1.0 + 1e-6 * gx[:, None] * (0.5 + gy[None, :] ** 2))
defs = {
    "Pressure": dict(kind="primary"), "Saturation": dict(kind="primary"), "PoreVolume": dict(kind="parameter"),
    "Mass": dict(kind="product", deps=["Density", "Saturation", "PoreVolume"], output=True),
    "Mobility": dict(kind="quotient", deps=["Kr", "Viscosity"], output=True),
    "OilMobility": dict(kind="quotient", deps=["KrO", "ShrinkageVisc"], c=[2.0], output=True),
    "KrO": dict(kind="power", deps=["So"], c=[0.8, 2.5, 0.15, 0.7]),
    "So": dict(kind="affine", deps=["Saturation"], c=[1.0, -1.0]),
    "Kr": dict(kind="power", deps=["Saturation"], c=[0.9, 2.0, 0.1, 0.8]),
    "Viscosity": dict(kind="table1d", deps=["Pressure"], table=t1),
    "ShrinkageVisc": dict(kind="table2d", deps=["Pressure", "So"], c=[1e-3], table=t2),
    "Density": dict(kind="exp", deps=["Pressure"], c=[1000.0, 4.5e-10, 1e5], output=True),
    "OilDensity": dict(kind="exp", deps=["Pressure"], c=[700.0, 1e-9, 1e5], output=True),
}
sv = J.SecondaryVariables(ctx, nc, defs)
state = {"Pressure": ctx.transfer(rng.uniform(5e4, 4e5, nc)), "Saturation": ctx.transfer(rng.uniform(0, 1, nc)), "PoreVolume": ctx.transfer(rng.uniform(1, 2, nc))}
out = {n: ctx.zeros(3 * nc) for n in sv.outputs}
ms = timed(lambda: sv.update_secondary_variables(state, out))
line("varprog_kernel<2> (10 secondary variables, 5 outputs with partials)", ms, nc * (3 * 8 + len(sv.outputs) * 3 * 8), cells=nc)
f = ctx.zeros(nc); df = ctx.zeros(nc)
ms = timed(lambda: t1.interpolate(state["Pressure"], f, df))
line("table_eval1_kernel (LinearInterpolant, 66 nodes, value + slope)", ms, nc * 24, points=nc)
for a in list(out.values()) + [f, df]:
    a.free()
del sv

# ---- two-phase system: assembly variants, transpose, Schur correction
sim = J.TwoPhaseSimulator(ctx, w["N"], nc, w["Tf"], w["gdz"], w["pv"], w["params"], ordering="multicolor")
sim.set_forces(w["src_cells"], w["src_vals"]); sim.set_state(w["p0"], w["sw0"])
alg = J.workloads.algorithmic_bytes(nc, nf)
for variant in ("cells", "faces"):
    ms = timed(lambda: sim.law.update_equation_and_linearized_system(sim.p, sim.s, sim.M0, w["dt"], sim.r, variant=variant), reps=5)
    line("two-phase assembly, variant %s (%s)" % (variant, "row owner, no atomics" if variant == "cells" else "fvm_face_assembly!: face scatter, warp-aggregated FP64 atomics"),
         ms, alg["assembly"], cells=nc, note="includes the secondary-variable (density) kernel")
# the same law assembled from the variable graph's output planes (closed forms of the built-in law as graph)
par = w["params"]
gdefs = {
    "Pressure": dict(kind="primary"), "Sw": dict(kind="primary"), "PoreVolume": dict(kind="parameter"),
    "So": dict(kind="affine", deps=["Sw"], c=[1.0, -1.0]),
    "DensityW": dict(kind="exp", deps=["Pressure"], c=[par[0], par[2], par[6]], output=True),
    "DensityO": dict(kind="exp", deps=["Pressure"], c=[par[1], par[3], par[6]], output=True),
    "KrW": dict(kind="product", deps=["Sw", "Sw"]), "KrO": dict(kind="product", deps=["So", "So"]),
    "MobilityW": dict(kind="product", deps=["DensityW", "KrW"], c=[1.0 / par[4]], output=True),
    "MobilityO": dict(kind="product", deps=["DensityO", "KrO"], c=[1.0 / par[5]], output=True),
    "MassW": dict(kind="product", deps=["PoreVolume", "DensityW", "Sw"], output=True),
    "MassO": dict(kind="product", deps=["PoreVolume", "DensityO", "So"], output=True),
}
sv2 = J.SecondaryVariables(ctx, nc, gdefs)
sw_dev = ctx.transfer(w["sw0"] if sim.perm is None else sim._cells_to_device_host(w["sw0"]))
pv_dev = ctx.transfer(w["pv"] if sim.perm is None else sim._cells_to_device_host(w["pv"]))
gstate = {"Pressure": sim.p, "Sw": sw_dev, "PoreVolume": pv_dev}
gout = {n: ctx.zeros(3 * nc) for n in sv2.outputs}
ms = timed(lambda: sv2.update_secondary_variables(gstate, gout), reps=5)
line("varprog_kernel<2> (two-phase property graph: 9 secondary variables, 6 outputs with partials)", ms, nc * (3 * 8 + 6 * 3 * 8), cells=nc)
ms = timed(lambda: J.assemble_with_properties(sim.law, sim.p, gout, sim.M0, w["dt"], sim.r), reps=5)
line("two-phase assembly from property planes (props_pack_kernel + twophase_props_kernel)", ms, alg["assembly"] + nc * (13 * 8 + 128) + nc * 18 * 8, cells=nc,
     note="bytes: SURVEY formula + planes read (18 doubles per cell) + 128-byte records written and read")
for a in gout.values():
    a.free()
nb = sim.jac.nnz
JT = sim.jac.adjoint()
ms = timed(lambda: JT.update_adjoint(), reps=5)
line("csr_transpose_kernel<2> (J^T for adjoint solves)", ms, nb * (64 + 4), blocks=nb)
x = ctx.transfer(rng.standard_normal(2 * nc)); y = ctx.zeros(2 * nc)
ms0 = timed(lambda: sim.jac.mul(y, x), reps=10)
line("jb_spmv (all rows)", ms0, alg["spmv"], blocks=nb)
sizes = [64] * 16        # 16 wells of 64 dofs
Cp, Dp, Ep, Cv, Dv, Ev = [], [], [], [], [], []
for m in sizes:
    rows = rng.integers(1, 2 * nc + 1, 4 * m); cols = rng.integers(1, m + 1, 4 * m)
    Cp.append((rows, cols, (2 * nc, m))); Cv.append(rng.standard_normal(4 * m))
    Dp.append((cols.copy(), rows.copy(), (m, 2 * nc))); Dv.append(rng.standard_normal(4 * m))
    ii, jj = np.meshgrid(np.arange(1, m + 1), np.arange(1, m + 1), indexing="ij")
    Ep.append((ii.ravel(), jj.ravel(), m)); Ev.append((np.eye(m) * 10 + rng.standard_normal((m, m))).ravel())
S = J.MultiLinearizedSystemSchur(sim.jac, Cp, Dp, Ep)
S.update(Cv, Dv, Ev)
ms1 = timed(lambda: S.mul(y, x), reps=10)
print(json.dumps(dict(kernel="schur_mul! = jb_spmv + schur_Dx / schur_Einv / schur_Cu (16 eliminated groups of 64 dofs)", ms=round(ms1, 4),
                      spmv_alone_ms=round(ms0, 4), correction_ms=round(ms1 - ms0, 4))), flush=True)

# ---- NFVM law: hex topology of 100^3 cells, 4 MPFA points per half discretisation taken from a window of nearby cell numbers
del S, JT, sim, x, y
w = J.workloads.unstructured_hex(100, 100, 100, permute=False)
nc, nf = w["nc"], w["nf"]
left = w["N"][:, 0].copy(); right = w["N"][:, 1].copy()


def half():
    ptr = 1 + 4 * np.arange(nf + 1, dtype=np.int64)
    near = np.clip(np.repeat(left, 4) + rng.integers(-120, 121, 4 * nf), 1, nc).astype(np.int64)
    return dict(T_left=rng.uniform(0.5, 2.0, nf), T_right=-rng.uniform(0.5, 2.0, nf), ptr=ptr, cell=near, T=rng.normal(0, 0.1, 4 * nf))


disc = J.NFVMDiscretization(ctx, left, right, nc, half(), half(), scheme="ntpfa")
jac = disc.declare_pattern()
disc.align_to_jacobian(jac)
p = ctx.transfer(rng.uniform(1, 2, nc)); r = ctx.zeros(nc)
nslots = disc.stencil()[1].shape[0]
ms = timed(lambda: disc.update_equation_and_linearized_system(p, r), reps=5)
line(":fvm assembly of the NFVM law (ntpfa, 1M cells, up to 10 stencil cells per face)", ms, nslots * (32 + 4 + 8 + 16) + jac.nnz * 8 + nc * 16, faces=nf, stencil_slots=int(nslots), nnz=int(jac.nnz),
     note="bytes: coefficients + cell id + positions + gathered p per slot, Jacobian zeroed + written, r; atomics hit L2")
