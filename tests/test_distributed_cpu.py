"""CPU, world_size 2 over gloo: the host side of the multi-GPU path -- rank/size plumbing, the
deterministic deal of grid batches over ranks, the id broadcast that replaces the reference's
MPI_Bcast of the ncclUniqueId, and sum-over-ranks == single-rank result (partials computed by the
oracle here; on GPUs the same sum is the NCCL allreduce of bench.py --gpus N)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
    import pyoracle as orc
    from gauxc_b200 import systems
    from gauxc_b200.driver import System, broadcast_bytes, dist_env
    rank, local_rank, world = dist_env()
    dist.init_process_group("gloo")
    assert (rank, world) == (dist.get_rank(), dist.get_world_size()) and world == 2
    # the unique-id broadcast path (any 128 bytes)
    payload = bytes(range(128)) if rank == 0 else b"\\0" * 128
    assert broadcast_bytes(payload, 128, 0) == bytes(range(128))
    orc.init_blas()
    def partial(rank, size):
        s = System("water", rank=rank, size=size, device=False, func="PBE", grid="FineGrid")
        assert s.rt.comm_rank() == rank and s.rt.comm_size() == size
        t = s.lb.export_tasks()
        coords = np.array([a[1:] for a in s.atoms])
        t["weights"] = orc.ssf_weights(coords, t["npts"], t["iParent"], t["dist_nearest"], t["points"], t["weights"])
        r = orc.exc_vxc(s.basis.flat(), s.nbf, s.P, t, "PBE")
        return r, int(t["npts"].sum())
    r, npts = partial(rank, world)
    buf = torch.from_numpy(np.concatenate([r["vxc"].ravel(), [r["exc"], r["nel"], float(npts)]]))
    dist.all_reduce(buf)
    if rank == 0:
        w, npts_w = partial(0, 1)
        got = buf.numpy()
        assert int(got[-1]) == npts_w, (got[-1], npts_w)
        assert abs(got[-3] - w["exc"]) < 1e-11 and abs(got[-2] - w["nel"]) < 1e-11
        assert np.abs(got[:-3] - w["vxc"].ravel()).max() < 1e-11
        print("DIST_OK")
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_partition_and_reduction_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout
