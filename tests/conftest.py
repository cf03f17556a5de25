import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library and the oracle must exist; build them if they are missing."""
    from gauxc_b200 import build as b
    b.build_library()
    b.build_driver()
    b.build_oracle()


@pytest.fixture(scope="session")
def orc():
    import pyoracle
    pyoracle.init_blas()
    return pyoracle


def make_lb(atoms, shells, grid="UltraFineGrid", pruning="Unpruned", normalize=True, device=False,
            rank=0, size=1, batch=512):
    import gauxc_b200 as gx
    mol = gx.Molecule(atoms)
    basis = gx.BasisSet(shells, normalize=normalize)
    mg = gx.MolGrid(mol, pruning, batch, "MuraKnowles", grid)
    rt = gx.RuntimeEnvironment(rank=rank, size=size, device=device)
    lb = gx.LoadBalancerFactory("Host", "Replicated").get_instance(rt, mol, mg, basis)
    return mol, basis, lb


@pytest.fixture(scope="session")
def benzene_golden():
    """Per golden file: system + tasks of the product's load balancer at tol = eps (the
    reference's test setting, tests/xc_integrator.cxx:163-167)."""
    from gauxc_b200 import systems

    cache = {}

    def get(name, pruning="Unpruned"):
        key = (name, pruning)
        if key not in cache:
            atoms, shells, P, VXC, EXC = systems.golden_system(name)
            for s in shells:
                s["tol"] = np.finfo(float).eps
            cache[key] = (atoms, shells, P, VXC, EXC)
        return cache[key]

    return get
