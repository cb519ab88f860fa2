import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def J():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def O():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx(J):
    c = J.B200Context(0)
    yield c
    c.close()


def oracle_system(O, w, bs=2):
    """Reference-side setup chain on a workload dict: half-face map -> pattern -> alignment."""
    hf = O.half_face_map(w["N"], w["nc"])
    I, Jc = O.tpfa_pattern(hf)
    rowptr, colidx = O.csr_from_coo(I, Jc, w["nc"])
    dpos, hpos = O.tpfa_alignment(hf, rowptr, colidx)
    return dict(hf=hf, rowptr=rowptr, colidx=colidx, diag_pos=dpos, hf_pos=hpos)


def to_scipy(n, bs, rowptr, colidx, nz):
    import scipy.sparse as sp
    blocks = np.asarray(nz).reshape(-1, bs, bs).transpose(0, 2, 1)   # column-major blocks -> [row, col]
    return sp.bsr_matrix((blocks, np.asarray(colidx) - 1, np.asarray(rowptr) - 1), shape=(n * bs, n * bs)).tocsr()
