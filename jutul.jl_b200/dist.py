"""Multi-GPU Newton hot path: one process per GPU (torchrun), METIS partition, owned-first local numbering with one
ghost layer, block-Jacobi ILU(0) per rank, halo exchange + Krylov reductions over NCCL (inside libjutul_b200.so, on the
library's stream). Mirrors the reference's distributed backend (ext/JutulPartitionedArraysExt):

    partition_boundary / remap / distribute_case   utils.jl:9-56,91-148   -> decompose()
    PArraySimulator, perform_step!                  interface.jl:2-97, overloads.jl:155-237 -> DistTwoPhaseSimulator
    parray_synchronize_primary_variables            interface.jl:189-220  -> halo exchange of p, S after the update
    unit_diagonalize!                               linalg.jl:1-35        -> ghost rows = -I, r = 0 (set once: only owned rows are assembled)
    local ILU(0) with ghost input zeroed            linalg.jl:78-88       -> ILU(0) partition {owned, ghost}
    parray_linear_solve!                            krylov.jl:1-105       -> jb_krylov_set_dist

torch.distributed is plumbing only (rendezvous, NCCL id broadcast, barriers); the data path never touches it.
"""
import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib
from ._lib import check


def _pkg():
    import sys
    return sys.modules[__package__]


# ---------------------------------------------------------------------------------------------- decomposition (host, numpy)
def decompose(N, nc, part, rank, local_order="default"):
    """Local problem of `rank` (0-based) for the cell partition `part` (1-based labels).

    Local cells are [owned | ghost]: owned in ascending global index (the reference's `:default` order,
    interface.jl:4) or in multicolour order of the owned sub-graph; ghosts = cells of other ranks sharing a face with an
    owned cell, ordered by (owner rank, global index) so every neighbour's data is one contiguous range.
    Local faces = faces with at least one owned cell, in ascending global face index.
    Returns a dict with the maps and the halo plan (send lists / receive ranges per neighbour rank).
    """
    N = np.asarray(N, dtype=np.int64)
    part0 = np.asarray(part, dtype=np.int64) - 1
    l, r = N[:, 0] - 1, N[:, 1] - 1
    pl, pr = part0[l], part0[r]
    face_mask = (pl == rank) | (pr == rank)
    faces = np.nonzero(face_mask)[0]
    owned = np.nonzero(part0 == rank)[0]
    fl, fr = l[faces], r[faces]
    # ghosts: far end of faces that leave the rank
    out_l = pl[faces] != rank
    out_r = pr[faces] != rank
    ghost = np.unique(np.concatenate([fl[out_l], fr[out_r]]))
    gown = part0[ghost]
    order = np.lexsort((ghost, gown))
    ghost, gown = ghost[order], gown[order]
    neigh, recv_counts = np.unique(gown, return_counts=True)
    recv_ptr = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    n_owned, n_ghost = owned.shape[0], ghost.shape[0]
    # optional renumbering of the owned cells (multicolour order of the owned sub-graph)
    ncolors = None
    if local_order == "multicolor" and n_owned > 1:
        J = _pkg()
        inner = ~out_l & ~out_r
        g2o = np.full(nc, -1, dtype=np.int64); g2o[owned] = np.arange(n_owned)
        Nin = np.stack([g2o[fl[inner]], g2o[fr[inner]]], axis=1) + 1
        # owned cells that touch a ghost go to the end of their colour: interior rows become contiguous ranges whose SpMV
        # overlaps the halo exchange
        bnd = None
        # (the same grouping lets whole chunks of first-colour rows without a ghost coupling qualify as identity rows of
        # the right-preconditioned operator, csrc/krylov.cu)
        if os.environ.get("JB_OVERLAP", "0") == "1" or os.environ.get("JB_RB_IDENTITY", "1") != "0":
            bnd = np.zeros(n_owned, dtype=np.int64)
            bnd[g2o[fl[out_r]]] = 1; bnd[g2o[fr[out_l]]] = 1
        perm, ncolors = J.multicolor_ordering(Nin, n_owned, last=bnd)
        owned_sorted = np.empty(n_owned, dtype=np.int64)
        owned_sorted[perm - 1] = owned
        owned = owned_sorted
    cells = np.concatenate([owned, ghost])
    g2l = np.full(nc, -1, dtype=np.int64)
    g2l[cells] = np.arange(cells.shape[0])
    N_local = np.stack([g2l[fl], g2l[fr]], axis=1) + 1
    # send lists: owned cells that neighbour q holds as ghosts = owned cells sharing a face with a cell of q,
    # in ascending GLOBAL index (q orders its ghosts the same way)
    send_idx, send_ptr = [], [0]
    for q in neigh:
        a = fl[(pr[faces] == q) & ~out_l]
        b = fr[(pl[faces] == q) & ~out_r]
        s = np.unique(np.concatenate([a, b]))
        send_idx.append(g2l[s] + 1)
        send_ptr.append(send_ptr[-1] + s.shape[0])
    send_idx = np.concatenate(send_idx).astype(np.int64) if send_idx else np.zeros(0, dtype=np.int64)
    return dict(rank=rank, n_owned=n_owned, n_ghost=n_ghost, n_local=n_owned + n_ghost, cells=cells, owned=owned, ghost=ghost,
                faces=faces, N_local=N_local, neigh=neigh.astype(np.int32), recv_ptr=recv_ptr,
                send_ptr=np.asarray(send_ptr, dtype=np.int64), send_idx=send_idx, g2l=g2l, ncolors=ncolors)


def halo_exchange_host(plan, vec, bs, td):
    """Reference implementation of the exchange with torch.distributed point-to-point (gloo on CPU): used by the CPU tests
    to validate the plan that the NCCL path executes on the device."""
    import torch
    v = vec.reshape(-1, bs)
    reqs, recv_bufs = [], []
    for k, q in enumerate(plan["neigh"]):
        s0, s1 = plan["send_ptr"][k], plan["send_ptr"][k + 1]
        sb = torch.from_numpy(np.ascontiguousarray(v[plan["send_idx"][s0:s1] - 1]))
        reqs.append(td.isend(sb, int(q)))
        rb = torch.empty((int(plan["recv_ptr"][k + 1] - plan["recv_ptr"][k]), bs), dtype=torch.float64)
        recv_bufs.append(rb)
        reqs.append(td.irecv(rb, int(q)))
    for rq in reqs:
        rq.wait()
    for k, rb in enumerate(recv_bufs):
        a, b = plan["n_owned"] + plan["recv_ptr"][k], plan["n_owned"] + plan["recv_ptr"][k + 1]
        v[a:b] = rb.numpy()
    return vec


# ---------------------------------------------------------------------------------------------- device-side objects
class _StdoutToStderr:
    """NCCL prints its version banner on stdout at communicator creation; bench.py must print exactly one JSON line,
    so fd 1 is pointed at fd 2 while communicators are being created."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        import sys
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class Communicator:
    """NCCL communicator owned by libjutul_b200.so; the 128-byte id travels through torch.distributed."""

    def __init__(self, ctx, rank, world, td):
        import torch
        self.ctx, self.rank, self.world = ctx, rank, world
        with _StdoutToStderr():
            buf = C.create_string_buffer(128)
            if rank == 0:
                check(ctx.lib.jb_nccl_unique_id(buf), ctx.h, "jb_nccl_unique_id")
            t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            if td.get_backend() == "nccl":
                t = t.cuda(ctx.device); td.broadcast(t, 0); t = t.cpu()
            else:
                td.broadcast(t, 0)
            idb = bytes(t.numpy().tobytes())
            h = C.c_void_p()
            check(ctx.lib.jb_comm_create(ctx.h, rank, world, idb, C.byref(h)), ctx.h, "jb_comm_create")
            self.h = h
            ctx.synchronize()

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.jb_comm_destroy(self.h); self.h = None


class HaloPlan:
    def __init__(self, comm, plan, td=None):
        ctx = comm.ctx
        self.ctx, self.comm, self.plan = ctx, comm, plan
        neigh = np.ascontiguousarray(plan["neigh"], dtype=np.int32)
        sp = np.ascontiguousarray(plan["send_ptr"], dtype=np.int64)
        si = np.ascontiguousarray(plan["send_idx"], dtype=np.int64)
        rp = np.ascontiguousarray(plan["recv_ptr"], dtype=np.int64)
        h = C.c_void_p()
        check(ctx.lib.jb_dist_create(comm.h, plan["n_owned"], plan["n_local"], neigh.shape[0], neigh.ctypes.data_as(_lib.PI32),
                                     sp.ctypes.data_as(_lib.PI64), si.ctypes.data_as(_lib.PI64), rp.ctypes.data_as(_lib.PI64), C.byref(h)),
              ctx.h, "jb_dist_create")
        self.h = h
        self.p2p = False
        if td is not None and os.environ.get("JB_P2P", "1") != "0":
            self._setup_peer_memory(td)

    def _setup_peer_memory(self, td):
        """All-gather the IPC handles of the symmetric buffers and every rank's staging offsets (setup only). If ANY rank
        cannot export or map the peer buffers (world > 16, no peer access, a multi-node launch), all ranks agree on it with a
        reduction and continue on the NCCL path together; the collectives below are executed by every rank either way."""
        import torch
        ctx, comm, plan = self.ctx, self.comm, self.plan
        W = comm.world
        buf = C.create_string_buffer(64)
        ok = 1
        why = ""
        if ctx.lib.jb_dist_p2p_export(self.h, buf) != 0:
            ok, why = 0, "jb_dist_p2p_export: " + _lib.last_error(ctx.h)
        dev = torch.device("cuda", ctx.device) if td.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().to(dev)
        allh = [torch.zeros_like(mine) for _ in range(W)]
        td.all_gather(allh, mine)
        handles = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        # off[src] = where src's data starts in my staging / ghost section; off[W] = my ghost count, then my n_owned, n_local
        off = np.full(W + 4, -1, dtype=np.int64)
        for k, q in enumerate(plan["neigh"]):
            off[int(q)] = plan["recv_ptr"][k]
        off[W] = max(plan["n_ghost"], 1)
        off[W + 1] = plan["n_owned"]; off[W + 2] = plan["n_local"]; off[W + 3] = ok
        t_off = torch.from_numpy(off).to(dev)
        allo = [torch.zeros_like(t_off) for _ in range(W)]
        td.all_gather(allo, t_off)
        tab = np.stack([t.cpu().numpy() for t in allo])        # tab[q][src]
        if not np.all(tab[:, W + 3] == 1):
            ok = 0                                             # some rank could not export: nobody opens
        if ok:
            ro = np.array([tab[int(q)][comm.rank] for q in plan["neigh"]], dtype=np.int64)
            rc = np.array([tab[int(q)][W] for q in plan["neigh"]], dtype=np.int64)
            rno = np.array([tab[int(q)][W + 1] for q in plan["neigh"]], dtype=np.int64)
            rnl = np.array([tab[int(q)][W + 2] for q in plan["neigh"]], dtype=np.int64)
            assert np.all(ro >= 0), "halo plans of neighbouring ranks are inconsistent"
            if ctx.lib.jb_dist_p2p_open(self.h, handles, ro.ctypes.data_as(_lib.PI64), rc.ctypes.data_as(_lib.PI64),
                                        rno.ctypes.data_as(_lib.PI64), rnl.ctypes.data_as(_lib.PI64)) != 0:
                ok, why = 0, "jb_dist_p2p_open: " + _lib.last_error(ctx.h)
        ctx.synchronize()
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        td.all_reduce(flag, op=td.ReduceOp.MIN)               # every rank takes the same decision
        if int(flag.item()) == 1:
            self.p2p = True
        else:
            ctx.lib.jb_dist_p2p_disable(self.h)
            self.p2p = False
            if comm.rank == 0 or why:
                print(f"[jutul_b200] rank {comm.rank}: peer-memory collectives unavailable ({why or 'another rank failed'}); using NCCL", flush=True)

    def check(self):
        st = self.ctx.lib.jb_dist_p2p_status(self.h)
        if st != 0:
            raise _lib.JutulB200Error(f"peer-memory collective timed out (code {st}): a rank is gone or the ranks diverged")

    def exchange(self, vec, bs):
        check(self.ctx.lib.jb_dist_halo_exchange(self.h, vec.ptr, bs), self.ctx.h, "jb_dist_halo_exchange")

    def allreduce(self, vals, op="sum"):
        v = np.ascontiguousarray(vals, dtype=np.float64).copy()
        check(self.ctx.lib.jb_dist_allreduce(self.h, v.ctypes.data_as(_lib.PF64), v.shape[0], 1 if op == "max" else 0), self.ctx.h,
              "jb_dist_allreduce")
        return v


class DistTwoPhaseSimulator:
    """Per-rank simulator of the distributed two-phase problem (PArraySimulator analogue)."""

    def __init__(self, ctx, comm, w, part, rtol=1e-3, atol=None, max_linear_iterations=100, tolerance=1e-3, max_nonlinear_iterations=15,
                 dp_abs_max=None, ds_abs_max=0.2, local_order="default", td=None):
        J = _pkg()
        self.ctx, self.comm = ctx, comm
        self.plan = plan = decompose(w["N"], w["nc"], part, comm.rank, local_order)
        self.n_owned, self.n_local = plan["n_owned"], plan["n_local"]
        cells, faces = plan["cells"], plan["faces"]
        self.disc = J.TwoPointPotentialFlowHardCoded(ctx, plan["N_local"], self.n_local)
        self.jac = J.tpfa_jacobian(self.disc, 2)
        self.storage = J.ConservationLawTPFAStorage(self.disc, self.jac)
        self.law = J.TwoPhaseConservationLaw(self.storage, w["Tf"][faces], w["gdz"][faces], w["pv"][cells], w["params"])
        check(ctx.lib.jb_twophase_set_owned(self.law.h, self.n_owned), ctx.h, "jb_twophase_set_owned")
        self.halo = HaloPlan(comm, plan, td)
        lp = np.ones(self.n_local, dtype=np.int64); lp[self.n_owned:] = 2      # ghosts decoupled: local ILU(0) of the owned block
        self.prec = J.ILUZeroPreconditioner(self.jac, lp if plan["n_ghost"] > 0 else None)
        self.krylov = J.GenericKrylov(self.jac, "bicgstab", self.prec, relative_tolerance=rtol, absolute_tolerance=atol,
                                      max_iterations=max_linear_iterations)
        check(ctx.lib.jb_krylov_set_dist(self.krylov.h, self.halo.h), ctx.h, "jb_krylov_set_dist")
        self.tolerance, self.max_nonlinear_iterations = tolerance, max_nonlinear_iterations
        self.dp_abs_max, self.ds_abs_max = dp_abs_max, ds_abs_max
        nl = self.n_local
        self.p = ctx.empty(nl); self.s = ctx.empty(2 * nl); self.M0 = ctx.zeros(2 * nl)
        self.r = ctx.zeros(2 * nl); self.dx = ctx.zeros(2 * nl)
        # ghost rows: -I on the diagonal, r = 0 (unit_diagonalize!); never touched again because only owned rows are assembled
        self.jac.unit_diagonalize_ghosts(self.r, self.n_owned)
        # sources on owned cells only
        g2l = plan["g2l"]
        sc = np.asarray(w["src_cells"], dtype=np.int64) - 1
        mine = (np.asarray(part)[sc] - 1) == comm.rank
        if mine.any():
            self.law.apply_forces(g2l[sc[mine]] + 1, np.asarray(w["src_vals"])[mine])
        self.set_state(w["p0"][cells], w["sw0"][cells])

    def set_state(self, p_local, sw_local):
        s = np.stack([sw_local, 1.0 - sw_local], axis=1).ravel()
        self.p.set(p_local); self.s.set(s)
        self.update_before_step()

    def update_before_step(self):
        self.law.total_masses(self.p, self.s, self.M0)

    def owned_state(self):
        return self.p.get()[: self.n_owned], self.s.get().reshape(-1, 2)[: self.n_owned, 0].copy()

    def perform_step(self, dt, solve=True):
        J = _pkg()
        rep = {}
        self.law.update_equation_and_linearized_system(self.p, self.s, self.M0, dt, self.r)
        e_loc = J.convergence_criterion(self.ctx, self.r, 2, self.n_owned)
        e = self.halo.allreduce(np.where(np.isfinite(e_loc), e_loc, np.inf), "max")       # reduce(max, errors)
        rep["errors"] = e
        if not np.all(np.isfinite(e)):
            rep["failure"] = "non-finite residual"
            return False, e, rep
        converged = bool(np.all(e <= self.tolerance))
        if converged or not solve:
            return converged, e, rep
        # the refactorisation status is max-reduced so that all ranks abort together (a rank that skipped the solve would
        # stall the peers' collectives)
        ok, its, hist, st = J.linear_solve(self.krylov, self.r, self.dx,
                                           status_reduce=lambda v: self.halo.allreduce([float(v)], "max")[0])
        rep["linear_iterations"], rep["linear_status"], rep["linear_residuals"] = its, st, hist
        if not ok:
            rep["linear_warning"] = f"Linear solver: status {st} after {its} iterations"
        J.update_primary_variable(self.ctx, self.p, self.dx, self.n_owned, dx_stride=2, abs_max=self.dp_abs_max)
        J.unit_update_pairs(self.ctx, self.s, self.dx.offset(1), self.n_owned, dx_stride=2, abs_max=self.ds_abs_max)
        self.halo.exchange(self.p, 1); self.halo.exchange(self.s, 2)                        # parray_synchronize_primary_variables
        if self.halo.p2p:
            self.halo.check()
        return False, e, rep

    def solve_ministep(self, dt):
        reports = []
        for it in range(1, self.max_nonlinear_iterations + 2):
            converged, e, rep = self.perform_step(dt, solve=it <= self.max_nonlinear_iterations)
            reports.append(rep)
            if converged or "failure" in rep:
                return converged, reports
        return False, reports


# ---------------------------------------------------------------------------------------------- bench (N > 1)
def run_bench(args, J):
    import torch
    import torch.distributed as td
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    with _StdoutToStderr():
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        td.barrier()
    from bench import METRIC, UNIT, ClockSampler, measured_peak, run_timestep
    nx, ny, nz = args.dims
    w = J.workloads.unstructured_hex(nx, ny, nz)
    nc, nf = w["nc"], w["nf"]
    ctx = J.B200Context(local_rank)
    comm = Communicator(ctx, rank, world, td)
    # METIS k-way on rank 0 (transmissibility weights as src/partitioning.jl:84-88), broadcast like MPI.Bcast!(p) (interface.jl:185)
    part_t = torch.zeros(nc, dtype=torch.int64, device="cuda")
    t0 = time.perf_counter()
    if rank == 0:
        part_t.copy_(torch.from_numpy(J.partition(w["N"], world, weights=w["Tf"], nc=nc)))
    td.broadcast(part_t, 0)
    part = part_t.cpu().numpy()
    t_part = time.perf_counter() - t0
    sim = DistTwoPhaseSimulator(ctx, comm, w, part, rtol=args.rtol, max_linear_iterations=args.max_linear_iterations, tolerance=args.tolerance,
                                local_order="multicolor" if args.ordering == "multicolor" else "default", td=td)
    dt = w["dt"]
    p_init, s_init = ctx.transfer(sim.p.get()), ctx.transfer(sim.s.get())

    unconverged = [0]

    def step(solve=True):
        conv, e, rep = sim.perform_step(dt, solve=solve)
        if "linear_warning" in rep:
            unconverged[0] += 1
        return conv, rep.get("linear_iterations", 0)

    def timestep():
        """One implicit timestep; the simulation advances (state0 <- state) unless --fixed-state (bench.py run_b200)."""
        if args.fixed_state:
            sim.p.copy_from(p_init); sim.s.copy_from(s_init)
            return run_timestep(step)
        res = run_timestep(step)           # every rank takes the same decisions: errors and iteration counts are all-reduced
        if res[0]:
            sim.update_before_step()
        else:
            sim.p.copy_from(p_init); sim.s.copy_from(s_init)
        p_init.copy_from(sim.p); s_init.copy_from(sim.s)
        return res

    def mark():
        return ctx.transfer(sim.p.get()), ctx.transfer(sim.s.get())

    def rewind(saved):
        sim.p.copy_from(saved[0]); sim.s.copy_from(saved[1]); p_init.copy_from(saved[0]); s_init.copy_from(saved[1])
        sim.update_before_step()

    def barrier():
        ctx.synchronize(); td.barrier(); torch.cuda.synchronize()

    for _ in range(args.warmup):
        timestep()
    start_state = mark()
    unconverged[0] = 0
    launches0 = ctx.launch_count
    clocks = ClockSampler(local_rank); clocks.start()
    barrier()
    J.timer_start(ctx)
    results = [timestep() for _ in range(args.steps)]
    ms = J.timer_stop(ctx)
    barrier()
    clk = clocks.stop()
    launches = ctx.launch_count - launches0
    mt = torch.tensor([ms], dtype=torch.float64, device="cuda")
    td.all_reduce(mt, op=td.ReduceOp.MAX)
    ms = float(mt.item())
    n_newton = sum(r[1] for r in results)
    lin_its = [x for r in results for x in r[2]]
    value = n_newton / (ms * 1e-3)

    # roofline pass (per rank; rank 0 reports its own kernels, algorithmic bytes of its local problem)
    nb_loc = sim.jac.nnz
    alg = J.workloads.algorithmic_bytes(sim.n_local, (nb_loc - sim.n_local) // 2, 2)
    peak, peak_kind = measured_peak()
    n_unconverged = unconverged[0]
    rewind(start_state)
    with J.DeviceProfile(ctx) as prof:
        barrier(); J.timer_start(ctx)
        for _ in range(max(1, min(args.steps, 2))):
            timestep()
        ms_prof = J.timer_stop(ctx)
        cls = prof.collect()
    kernels = {}
    bytes_of = {"assembly": alg["assembly"], "spmv": alg["spmv"], "ilu_apply": alg["ilu_apply"], "ilu_factor": alg["ilu_factor"]}
    id_chunks, id_rows, id_blocks = sim.krylov.identity_info()      # identity rows of A*N^-1 (csrc/krylov.cu): fewer blocks streamed
    if id_rows:
        bytes_of["spmv"] = (nb_loc - id_blocks) * 36 + 4 * (sim.n_local - id_rows + 1) + 2 * sim.n_local * 16 + id_rows * 16
    for k, bts in bytes_of.items():
        t, c = cls[k]
        if c:
            gbs = bts * c / (t * 1e-3) / 1e9
            kernels[k] = {"ms_per_launch": t / c, "launches": c, "achieved_gbs": gbs, "frac": gbs / peak, "share_of_step": t / ms_prof}
    for k in ("vector", "other"):
        t, c = cls[k]
        kernels["halo+allreduce" if k == "other" else k] = {"ms_total": t, "launches": c, "share_of_step": t / ms_prof}
    from bench import fused_kernel_report
    fused = fused_kernel_report(J, sim, cls, ms_prof, nb_loc, sim.n_local, peak)
    if fused:
        kernels["bicgstab_iteration"] = fused
        roofline = {"bound": "hbm", "kernel": fused["kernel"], "achieved": fused["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": fused["frac"], "traffic": None, "peak_source": peak_kind,
                    "algorithmic_bytes_per_launch": fused["algorithmic_bytes_per_launch"], "ms_per_launch": fused["ms_per_launch"],
                    "note": "rank 0, local sub-domain; collectives run inside this kernel"}
    else:
        dom = max(bytes_of, key=lambda k: cls[k][0])
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac"], "traffic": None, "peak_source": peak_kind, "algorithmic_bytes_per_launch": bytes_of[dom],
                    "note": "rank 0, local sub-domain"}

    # e2e: host state buffers of the owned cells go in and come back every Newton iteration
    rewind(start_state)
    p0_h, s0_h = p_init.get(), s_init.get()
    nloc = sim.n_local
    p_h, s_h = ctx.pinned_empty(nloc), ctx.pinned_empty(2 * nloc)      # pinned host state buffers
    p_h[:] = p0_h; s_h[:] = s0_h
    h2d = d2h = 0

    def step_host(solve=True):
        nonlocal h2d, d2h
        sim.p.set(p_h); sim.s.set(s_h); h2d += nloc * 3 * 8
        conv, e, rep = sim.perform_step(dt, solve=solve)
        if not conv:
            sim.p.get_into(p_h); sim.s.get_into(s_h); d2h += nloc * 3 * 8
        return conv, rep.get("linear_iterations", 0)

    def timestep_host():
        if args.fixed_state:
            p_h[:] = p0_h; s_h[:] = s0_h
            return run_timestep(step_host)
        res = run_timestep(step_host)
        sim.p.set(p_h); sim.s.set(s_h); sim.update_before_step()       # the host owns the state: state0 <- state
        return res

    barrier()
    t0 = time.perf_counter()
    res2 = [timestep_host() for _ in range(max(1, min(args.steps, 2)))]
    barrier()
    wall2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    td.all_reduce(wall2, op=td.ReduceOp.MAX)
    nn2 = sum(r[1] for r in res2)
    h2d_t = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda"); td.all_reduce(h2d_t)
    e2e = {"value": nn2 / float(wall2.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d_t[0].item() / len(res2)),
           "d2h_bytes_per_step": int(h2d_t[1].item() / len(res2)), "ms_per_step": 1e3 * float(wall2.item()) / len(res2),
           "api": "DistTwoPhaseSimulator.perform_step with host state buffers per rank"}
    sizes = torch.tensor([sim.n_owned, sim.plan["n_ghost"]], dtype=torch.int64, device="cuda")
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    td.all_gather(all_sizes, sizes)
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "b200",
            "config": {"workload": f"one implicit timestep (dt = 1 day) of the two-phase TPFA problem solved to Newton convergence from the "
                                   f"SURVEY §8(d) initial state, {nx}x{ny}x{nz} = {nc} cells, unstructured (permuted) hex grid, METIS k-way over "
                                   f"{world} GPUs, block-Jacobi ILU(0)-BiCGStab rtol={args.rtol:g}",
                       "cells": nc, "faces": nf, "block_size": 2, "linear_rtol": args.rtol, "max_linear_iterations": args.max_linear_iterations,
                       "newton_tolerance": args.tolerance, "parallelism": f"domain decomposition, {world} ranks (one process per GPU); halo exchange and Krylov all-reduce as peer-memory "
                                      f"stores over NVLink inside the fused iteration kernel (NCCL: fallback path and setup only)",
                       "l2_policy": "inputs larger than L2 per rank" if (nb_loc * 32) > 126e6 else "local Jacobian fits L2: no flush (strong scaling)",
                       "owned_ghost_per_rank": [[int(a[0]), int(a[1])] for a in all_sizes], "partition_seconds": t_part,
                       "cell_ordering": args.ordering, "ilu": sim.prec.info(), "operator_identity_rows_rank0": id_rows,
                       "collectives": "peer memory over NVLink (fused all-reduce + recurrence kernel, direct halo stores)" if sim.halo.p2p else "NCCL"},
            "newton_iterations_per_step": n_newton / max(args.steps, 1), "converged": all(r[0] for r in results),
            "linear_iterations_per_newton": float(np.mean(lin_its)) if lin_its else None, "linear_iterations": results[0][2],
            "linear_solves": len(lin_its), "linear_solves_unconverged": n_unconverged,
            "timestepping": "fixed state (every step re-solves report step 1)" if args.fixed_state else
                            f"advancing: warm-up = report steps 1..{args.warmup}, timed = report steps {args.warmup + 1}..{args.warmup + args.steps}",
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline, "kernels": kernels,
            "hbm_gbs": {k: kernels[k]["achieved_gbs"] for k in ("assembly", "spmv") if k in kernels}, "cpu_baseline": None,
        }
        print(json.dumps(out))
    barrier()
    td.destroy_process_group()
