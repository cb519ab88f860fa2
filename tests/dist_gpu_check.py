"""Launched by torchrun (N >= 2 GPUs): the distributed Newton step must reproduce the single-GPU run whose ILU(0) uses the
same METIS partition as block-Jacobi blocks: same Newton/linear iteration counts (+-), same iterate."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

J = g.load_package()
from jutul_b200 import dist as D  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
td.init_process_group("nccl", device_id=torch.device("cuda", lr))
dims = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "24,20,16").split(",")]
w = J.workloads.unstructured_hex(*dims)
nc = w["nc"]
ctx = J.B200Context(lr)
comm = D.Communicator(ctx, rank, world, td)
part = J.partition(w["N"], world, weights=w["Tf"], nc=nc)
ok_all = True
ref_cache = {}
for order, p2p, fused in (("default", "1", "1"), ("default", "0", "1"), ("multicolor", "1", "1"), ("multicolor", "1", "0"), ("multicolor", "0", "1")):
    if True:
        # fused = "1": the fused iteration kernel (peer-memory collectives inside the kernel) where the structure qualifies
        # (multicolour ordering, peer memory); "0": the multi-kernel driver on the same problem
        os.environ["JB_P2P"] = p2p; os.environ["JB_PERSISTENT"] = fused
        sim = D.DistTwoPhaseSimulator(ctx, comm, w, part, rtol=1e-9, tolerance=1e-6, max_linear_iterations=500, local_order=order, td=td)
        assert sim.halo.p2p == (p2p == "1")
        conv, reps = sim.solve_ministep(w["dt"])
        p_o, sw_o = sim.owned_state()
        # gather the owned states on every rank through torch.distributed (test harness only)
        pg = torch.zeros(nc, dtype=torch.float64, device="cuda"); sg = torch.zeros(nc, dtype=torch.float64, device="cuda")
        idx = torch.from_numpy(sim.plan["owned"]).cuda()
        pg[idx] = torch.from_numpy(p_o).cuda(); sg[idx] = torch.from_numpy(sw_o).cuda()
        td.all_reduce(pg); td.all_reduce(sg)
        if rank == 0:
            if "ref" not in ref_cache:
                ref = J.TwoPhaseSimulator(ctx, w["N"], nc, w["Tf"], w["gdz"], w["pv"], w["params"], partition=part, rtol=1e-9, tolerance=1e-6,
                                          max_linear_iterations=500)
                ref.set_forces(w["src_cells"], w["src_vals"])
                ref.set_state(w["p0"], w["sw0"])
                conv1, reps1 = ref.solve_ministep(w["dt"])
                ref_cache["ref"] = (conv1, reps1) + ref.get_state()
            conv1, reps1, p1, sw1 = ref_cache["ref"]
            ep = np.abs(pg.cpu().numpy() - p1).max() / np.abs(p1).max(); es = np.abs(sg.cpu().numpy() - sw1).max()
            its = [r.get("linear_iterations") for r in reps if "linear_iterations" in r]
            its1 = [r.get("linear_iterations") for r in reps1 if "linear_iterations" in r]
            good = conv and conv1 and len(reps) == len(reps1) and ep <= 1e-8 and es <= 1e-8
            if order == "default":
                # same elimination order inside every block => same arithmetic up to the summation order of the inner products.
                # BiCGStab amplifies that rounding over a long solve (rtol 1e-9: 150-350 iterations), so the counts agree to
                # ~15 % while the converged iterates agree to 1e-12: measured 264 vs 312 at 60^3 with dp = 8e-13
                # (profiles/r02b_dist_check_2gpu.log).
                good = good and all(abs(a - b) <= max(2, b // 4) for a, b in zip(its, its1))
            print(f"[dist_gpu_check] world={world} order={order} p2p={p2p} fused={fused} newton {len(reps)} vs {len(reps1)} lin {its} vs {its1} "
                  f"dp={ep:.2e} ds={es:.2e} -> {'OK' if good else 'MISMATCH'}", flush=True)
            ok_all = ok_all and good
        del sim
flag = torch.tensor([1 if ok_all else 0], device="cuda"); td.broadcast(flag, 0)
td.barrier()
td.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
