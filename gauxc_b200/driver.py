"""Host-side driver that mirrors the reference's tests/standalone_driver.cxx:31-477 over the
C ABI: Molecule -> MolGrid -> BasisSet -> LoadBalancer -> MolecularWeights(Device, SSF) ->
XCIntegrator(Device) for the BASELINE.json configurations, plus the one-process-per-GPU
bootstrap (rank/size from the launcher, NCCL id broadcast through torch.distributed instead of
the reference's MPI_Bcast, src/reduction_driver/device/nccl_reduction_driver.hpp:24-34).
"""
import os
import time

import numpy as np

from . import capi, systems


def dist_env():
    """(rank, local_rank, world_size) from torchrun's environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def broadcast_bytes(payload, nbytes, src=0):
    """Broadcast a byte string from rank `src` over the default torch.distributed group
    (works with gloo on CPU and nccl on GPU)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def init_nccl_from_torch():
    """Create the library's own NCCL communicator (used by the NCCL ReductionDriver) with the
    unique id broadcast through the already initialised torch.distributed group."""
    import torch.distributed as dist
    rank, size = dist.get_rank(), dist.get_world_size()
    uid = capi.nccl_get_unique_id() if rank == 0 else b"\0" * 128
    uid = broadcast_bytes(uid, 128, 0)
    capi.nccl_init(uid, rank, size)


class System:
    """One BASELINE.json configuration, set up through the C ABI."""

    def __init__(self, workload, rank=0, size=1, device=True, func=None, grid=None, pruning="Unpruned",
                 batch=512, basis_tol=1e-10, P=None, verbose=False):
        t0 = time.time()
        cfg = dict(systems.CONFIGS[workload])
        self.workload = workload
        self.func_name = func or cfg["func"]
        self.grid = grid or cfg["grid"]
        self.basis_name = cfg["basis"]
        if workload == "benzene":
            # geometry, basis and the real SCF density of the reference's PBE0 fixture
            atoms, shells, Pg, _, _ = systems.golden_system("benzene_pbe0_cc-pvdz_ufg_ssf")
            for s in shells:
                s["tol"] = basis_tol
            self.atoms, self.shells, self.P = atoms, shells, np.asfortranarray(Pg)
            normalize = False
        else:
            self.atoms = systems.config_atoms(workload)
            self.shells = systems.make_basis_shells(self.atoms, cfg["basis"], spherical=True, tol=basis_tol)
            self.P = None
            normalize = True
        import gauxc_b200 as gx
        self.mol = gx.Molecule(self.atoms)
        self.basis = gx.BasisSet(self.shells, normalize=normalize)
        self.nbf = self.basis.nbf()
        self.mg = gx.MolGrid(self.mol, pruning, batch, "MuraKnowles", self.grid)
        self.rt = gx.RuntimeEnvironment(rank=rank, size=size, device=device)
        # Device execution space: shell screening on the GPU (bit-identical task list, tests/test_gpu_parity.py)
        self.lb = gx.LoadBalancerFactory("Device" if device else "Host", "Replicated") \
            .get_instance(self.rt, self.mol, self.mg, self.basis)
        self.npts_local = self.lb.total_npts()
        self.t_setup = time.time() - t0
        if P is not None:
            self.P = np.asfortranarray(P)
        elif self.P is None:
            self.P = systems.synthetic_density(self.atoms, self.shells)
        self.t_density = time.time() - t0 - self.t_setup
        self.mw = None
        self.integrator = None
        if verbose:
            print(f"[{workload}] natoms={len(self.atoms)} nbf={self.nbf} local npts={self.npts_local} "
                  f"ntasks={self.lb.ntasks()} setup {self.t_setup:.1f}s density {self.t_density:.1f}s", flush=True)

    def modify_weights(self):
        import gauxc_b200 as gx
        self.mw = gx.MolecularWeightsFactory("Device", "Default", "SSF").get_instance()
        t0 = time.time()
        self.mw.modify_weights(self.lb)
        self.t_weights = time.time() - t0
        return self.mw.last_ms()

    def make_integrator(self, reduction="Default"):
        import gauxc_b200 as gx
        self.func = gx.Functional(self.func_name)
        self.integrator = gx.XCIntegratorFactory("Device", "Replicated", "Default", "Default", reduction) \
            .get_instance(self.func, self.lb)
        return self.integrator
