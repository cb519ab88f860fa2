"""Host-side logic of the multi-GPU path on CPU: decomposition maps and the halo plan, exercised with
world_size = 2 and 3 over gloo (the device path runs the same plan through NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, order, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as g
        J = g.load_package()
        from jutul_b200 import dist as D
        w = J.workloads.unstructured_hex(7, 6, 5)
        nc = w["nc"]
        part = J.partition(w["N"], world, weights=w["Tf"], nc=nc)
        # every rank computes the same METIS partition (deterministic), as if broadcast from rank 0
        pt = torch.from_numpy(part.copy()); td.broadcast(pt, 0)
        assert np.array_equal(pt.numpy(), part)
        plan = D.decompose(w["N"], nc, part, rank, order)
        no, nl = plan["n_owned"], plan["n_local"]
        assert sorted(plan["owned"].tolist()) == np.nonzero(part - 1 == rank)[0].tolist()
        assert np.all(part[plan["ghost"]] - 1 != rank)
        gown = part[plan["ghost"]] - 1
        assert np.all(np.diff(gown) >= 0)                       # ghosts grouped by owner rank
        for k, qn in enumerate(plan["neigh"]):
            seg = plan["ghost"][plan["recv_ptr"][k]:plan["recv_ptr"][k + 1]]
            assert np.all(gown[plan["recv_ptr"][k]:plan["recv_ptr"][k + 1]] == qn) and np.all(np.diff(seg) > 0)
        # local faces: both ends local, at least one owned; every face touching an owned cell is present
        NL = plan["N_local"]
        assert NL.min() >= 1 and NL.max() <= nl and np.all((NL[:, 0] <= no) | (NL[:, 1] <= no))
        gl = plan["cells"][NL - 1] + 1
        assert np.array_equal(gl, w["N"][plan["faces"]])
        touching = (part[w["N"][:, 0] - 1] - 1 == rank) | (part[w["N"][:, 1] - 1] - 1 == rank)
        assert np.array_equal(np.nonzero(touching)[0], plan["faces"])
        # halo exchange over gloo: ghosts receive the owner's values (bs = 1 and 2)
        for bs in (1, 2):
            glob = np.arange(nc * bs, dtype=np.float64).reshape(nc, bs) * 1.5 + 7
            loc = np.full((nl, bs), -1.0)
            loc[:no] = glob[plan["owned"]]
            D.halo_exchange_host(plan, loc.reshape(-1), bs, td)
            assert np.array_equal(loc, glob[plan["cells"]])
        # distributed SpMV == global SpMV on owned rows (ghost columns filled by the exchange)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle as O
        from conftest import oracle_system
        s = oracle_system(O, w)
        rng = np.random.default_rng(3)
        nzg = rng.standard_normal(s["colidx"].shape[0] * 4)
        xg = rng.standard_normal(2 * nc)
        yg = O.spmv(nc, 2, s["rowptr"], s["colidx"], nzg, xg).reshape(nc, 2)
        wl = dict(N=plan["N_local"], nc=nl)
        sl = oracle_system(O, wl)
        # local values: copy the global blocks of owned rows
        nzl = np.zeros(sl["colidx"].shape[0] * 4)
        for li in range(no):
            gi = plan["cells"][li]
            gcols = s["colidx"][s["rowptr"][gi] - 1:s["rowptr"][gi + 1] - 1] - 1
            for k in range(sl["rowptr"][li] - 1, sl["rowptr"][li + 1] - 1):
                gc = plan["cells"][sl["colidx"][k] - 1]
                kk = s["rowptr"][gi] - 1 + int(np.nonzero(gcols == gc)[0][0])
                nzl[4 * k:4 * k + 4] = nzg[4 * kk:4 * kk + 4]
        xl = np.zeros((nl, 2)); xl[:no] = xg.reshape(nc, 2)[plan["owned"]]
        D.halo_exchange_host(plan, xl.reshape(-1), 2, td)
        yl = O.spmv(nl, 2, sl["rowptr"], sl["colidx"], nzl, xl.reshape(-1)).reshape(nl, 2)
        assert np.allclose(yl[:no], yg[plan["owned"]], rtol=1e-13, atol=1e-13)
        # global inner product = all-reduce of owned-row partial sums
        t = torch.tensor([float(np.dot(xl[:no].ravel(), xl[:no].ravel()))], dtype=torch.float64)
        td.all_reduce(t)
        assert abs(t.item() - float(np.dot(xg, xg))) <= 1e-10 * float(np.dot(xg, xg))
        q.put((rank, "ok", no, plan["n_ghost"]))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world,order", [(2, "default"), (3, "default"), (2, "multicolor")])
def test_decomposition_and_halo_plan_gloo(world, order):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, order, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r[1]
    assert sum(r[2] for r in res) == 7 * 6 * 5
