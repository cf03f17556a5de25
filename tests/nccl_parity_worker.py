"""Worker of tests/test_gpu_parity.py::test_two_rank_nccl_result_matches_oracle (one process per GPU under
torchrun): every rank evaluates its share of the grid batches on its GPU, the library's NCCL reduction driver
sums VXC / EXC / N_el (and the EXC gradient), and EVERY rank checks the reduced result against the oracle on the
undivided task list."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import pyoracle as orc
    import gauxc_b200 as gx
    from gauxc_b200 import capi
    from gauxc_b200.driver import System, init_nccl_from_torch
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    capi.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    init_nccl_from_torch()
    orc.init_blas()
    TOL = 1e-10
    for workload, func, grid in (("benzene", "PBE", "FineGrid"), ("benzene", "SVWN5", "FineGrid"),
                                 ("water", "B3LYP", "FineGrid")):
        s = System(workload, rank=rank, size=world, device=True, func=func, grid=grid)
        s.modify_weights()
        integ = s.make_integrator("NCCL")
        exc, vxc = integ.eval_exc_vxc(s.P)
        nel = integ.stats()["n_el"]
        assert abs(integ.eval_exc(s.P) - exc) < 1e-12
        # the undivided problem on one rank, weights from the same Device kernel
        s1 = System(workload, rank=0, size=1, device=True, func=func, grid=grid)
        s1.modify_weights()
        ref = orc.exc_vxc(s1.basis.flat(), s1.nbf, s1.P, s1.lb.export_tasks(), func)
        d = (abs(exc - ref["exc"]), float(np.abs(vxc - ref["vxc"]).max()), abs(nel - ref["nel"]))
        print(f"[rank {rank}] {workload} {func}: local npts {s.npts_local} of {s1.npts_local}, "
              f"|dEXC| {d[0]:.2e} max|dVXC| {d[1]:.2e} |dNel| {d[2]:.2e}", flush=True)
        assert 0 < s.npts_local < s1.npts_local
        assert max(d) <= TOL, d
        assert np.array_equal(vxc, vxc.T)
        # EXC gradient: the 3 natoms sums are reduced on the device (the reference refuses a device reduction here)
        na = len(s.atoms)
        xyz = np.array([a[1:] for a in s.atoms])
        s2c = np.linalg.norm(s1.basis.flat()[5][:, None, :] - xyz[None, :, :], axis=2).argmin(1).astype(np.int32)
        for wd in (False, True):
            g = integ.eval_exc_grad(s.P, na, include_weight_derivatives=wd)
            go = orc.exc_grad(s1.basis.flat(), s2c, xyz, s1.nbf, s1.P, s1.lb.export_tasks(), func,
                              include_weight_derivatives=wd)
            dg = float(np.abs(g - go).max())
            print(f"[rank {rank}] {workload} {func}: EXC gradient (weight derivatives {wd}) max|dg| {dg:.2e}", flush=True)
            assert dg <= TOL, dg
        del integ, s, s1
    dist.barrier()
    if rank == 0:
        print("nccl parity ok", flush=True)
    capi.nccl_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
