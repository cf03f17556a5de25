#!/usr/bin/env python
"""bench.py -- EXC+VXC throughput of the B200 Device path (grid points/s), one process per GPU.

A "step" is one full XCIntegrator::eval_exc_vxc over the whole molecular grid of the workload
(all local grid batches: collocation -> X = P_sub B (DMMA) -> rho/grad rho -> functional -> Z ->
VXC += B^T Z (DMMA) -> scatter -> symmetrise -> allreduce over ranks).

  value   : grid points/s, whole job, inputs (P, tasks, weights) resident in HBM, device-resident
            entry point (gauxc_b200_integrator_eval_exc_vxc_rks_device), CUDA events on the
            integrator's stream, max over ranks.
  e2e     : same metric through the reference-facing C-ABI call gauxc_integrator_eval_exc_vxc_rks
            with HOST buffers (P from pinned host memory H2D, VXC + EXC D2H inside the timed region).
  roofline: the dominant kernel class (FP64 DMMA contractions), algorithmic flops / CUDA-event time.
  cpu_baseline: the oracle (CPU restatement of the reference Host path, OpenMP + OpenBLAS) timed
            on this box's host cores on a bounded task sample (N=1, rank 0 only).

`--impl reference` times the reference's CPU algorithm (oracle port; the reference itself cannot
be built here) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "exc_vxc_grid_points_per_s"
UNIT = "grid points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GAUXC_B200_WORKLOAD", "taxol"),
                    choices=["water", "benzene", "taxol", "ubiquitin", "water833"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w": float(np.median(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def oracle_sample(sys_, tasks, seconds, threads=None):
    """Time the oracle on every `stride`-th task, stride chosen for ~`seconds` of CPU work."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    blas = orc.init_blas()
    if threads:
        orc.lib().oracle_set_num_threads(int(threads))
    cores = orc.num_threads()
    fb = sys_.basis.flat()
    nt = len(tasks["npts"])
    cost = tasks["nbe"].astype(float) ** 2 * tasks["npts"] * 4
    # calibrate the stride on growing samples until one takes >= 1 s (first calls pay thread start-up)
    stride = max(1, nt // 64)
    while True:
        t0 = time.time()
        r = orc.exc_vxc(fb, sys_.nbf, sys_.P, tasks, sys_.func_name, task_stride=stride)
        dt = max(time.time() - t0, 1e-3)
        if dt >= 1.0 or stride == 1:
            break
        stride = max(1, stride // 4)
    rate = r["flops"] / dt  # dense flops/s seen by the calibration run
    stride = int(max(1, np.ceil(cost.sum() / max(rate * seconds, 1.0))))
    t0 = time.time()
    r = orc.exc_vxc(fb, sys_.nbf, sys_.P, tasks, sys_.func_name, task_stride=stride)
    dt = time.time() - t0
    npts = int(tasks["npts"][::stride].sum())
    return dict(value=npts / dt, unit=UNIT, cores=cores, kind="port", blas=os.path.basename(blas),
                sample=f"every {stride}th of {nt} tasks ({npts} of {int(tasks['npts'].sum())} points, "
                       f"{r['flops']:.3e} dense flops) in {dt:.2f} s",
                gflops=r["flops"] / dt / 1e9, seconds=dt, npts=npts)


def run_reference(args):
    """The reference's CPU algorithm (oracle port: OpenMP over tasks + OpenBLAS dgemm/dsyr2k, like
    reference_replicated_xc_host_integrator_exc_vxc.hpp:203-209) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gauxc_b200.driver import System
    s = System(args.workload, rank=0, size=1, device=False, verbose=args.verbose)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    tasks = s.lb.export_tasks()
    coords = np.array([a[1:] for a in s.atoms])
    # SSF weights belong to setup in both arms (MolecularWeights::modify_weights), not to the step
    stride_w = max(1, len(tasks["npts"]) // 2000)
    vals = []
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    base = None
    # all host threads this process may use (torchrun pins OMP_NUM_THREADS=1 for its workers)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for it in range(args.warmup + args.steps):
        base = oracle_sample(s, tasks, per_step, threads=ncores)
        if it >= args.warmup:
            vals.append(base)
    v = float(np.mean([b["value"] for b in vals]))
    ms = float(np.mean([b["seconds"] for b in vals])) * 1e3
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(s, 1, "host cores; bounded task sample per step"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": base["cores"], "kind": "port",
                             "sample": base["sample"], "blas": base["blas"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(s, ngpu, l2note):
    return {"workload": f"{s.workload} {s.basis_name} {s.func_name} {s.grid} unpruned SSF RKS EXC+VXC",
            "natoms": len(s.atoms), "nbf": int(s.nbf), "basis_tol": 1e-10, "batch_size": 512,
            "density": "SCF density of the reference fixture" if s.workload == "benzene" else
                       "synthetic SAD-like + seeded perturbation (SURVEY 8d)",
            "parallelism": f"grid batches dealt over {ngpu} GPU(s), NCCL allreduce of VXC/EXC",
            "l2": l2note}


def main():
    args = parse()
    # torchrun pins OMP_NUM_THREADS=1; the host-side setup (grid, screening, task merge -- outside the
    # timed region) and the CPU reference arm are OpenMP code: give every working rank its share of the
    # host cores instead (the reference arm runs on rank 0 alone, so it takes all of them)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if world_env > 1:
        share = ncores if args.impl == "reference" else max(1, ncores // world_env)
        os.environ["OMP_NUM_THREADS"] = str(share)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from gauxc_b200 import capi
    from gauxc_b200.driver import System, init_nccl_from_torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if capi.device_count() < 1 or not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the Device path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    capi.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        init_nccl_from_torch()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s = System(args.workload, rank=rank, size=world, device=True, verbose=args.verbose and rank == 0)
    ssf_ms = s.modify_weights()
    integ = s.make_integrator("Default")
    nbf = s.nbf
    info = s.lb.task_info()
    npts_local = int(info["npts"].sum())
    npts_t = torch.tensor([float(npts_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(npts_t)
    npts_total = int(npts_t.item())

    # ---- device-resident arm -------------------------------------------------------------------
    dP = torch.from_numpy(np.ascontiguousarray(s.P.T)).cuda()  # symmetric: layout agnostic
    dV = torch.zeros((nbf, nbf), dtype=torch.float64, device="cuda")
    dout = torch.zeros(2, dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(integ.stream())
    torch.cuda.synchronize()

    def step_dev():
        integ.eval_exc_vxc_device(dP.data_ptr(), dV.data_ptr(), dout.data_ptr())

    for _ in range(args.warmup):
        step_dev()
    integ.set_profile(True)  # per-kernel CUDA events on the launching stream, read after the sync
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = np.zeros(4)
    launches = 0
    e0.record(stream)
    t0 = time.time()
    for _ in range(args.steps):
        step_dev()
        st = integ.stats()
        kms += [st["k_colloc_ms"], st["k_xmat_ms"], st["k_zmat_ms"], st["k_vxc_ms"]]
        launches += int(st["launches"])
    e1.record(stream)
    barrier()
    wall_ms = (time.time() - t0) * 1e3
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    tt = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(tt[0]), float(tt[1])
    ms_per_step = dev_ms / args.steps
    value = npts_total / (ms_per_step * 1e-3)
    exc_dev, nel_dev = [float(x) for x in dout.cpu()]
    st = integ.stats()
    integ.set_profile(False)

    # ---- end-to-end arm: the reference-facing C-ABI call with host buffers -----------------------
    Ph = torch.empty((nbf, nbf), dtype=torch.float64).pin_memory()
    Ph.copy_(torch.from_numpy(np.ascontiguousarray(s.P.T)))
    Vh = torch.empty((nbf, nbf), dtype=torch.float64).pin_memory()
    Pn, Vn = Ph.numpy(), Vh.numpy()
    for _ in range(max(1, min(args.warmup, 2))):
        integ.eval_exc_vxc_raw(nbf, nbf, Pn, nbf, Vn, nbf)
    barrier()
    t0 = time.time()
    e2e_dev_ms = 0.0
    for _ in range(args.steps):
        exc_h = integ.eval_exc_vxc_raw(nbf, nbf, Pn, nbf, Vn, nbf)
        e2e_dev_ms += integ.stats()["total_ms"]
    barrier()
    e2e_wall_ms = (time.time() - t0) * 1e3
    tt = torch.tensor([e2e_wall_ms, e2e_dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms = float(tt[0]) / args.steps  # wall clock of the blocking host call, max over ranks
    e2e = {"value": npts_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "device_ms_per_step": float(tt[1]) / args.steps,
           "h2d_bytes_per_step": int(nbf * nbf * 8), "d2h_bytes_per_step": int(nbf * nbf * 8 + 16),
           "api": "gauxc_integrator_eval_exc_vxc_rks(host P, host VXC), pinned host buffers"}

    # ---- roofline of the dominant kernel class (local rank 0 numbers) ------------------------------
    peaks, peak_src = measured_peaks()
    dmma_peak = capi.probe_peak("dmma")  # FP64 tensor peak is not in MEASURED_PEAKS.json: probe
    f_dense = st["f_dense"]
    k_names = ["collocation", "xmat_density(DMMA)", "func_zmat", "vxc(DMMA)"]
    k_ms = (kms / args.steps).tolist()
    nb = max(1, int(st["nbatches"]))
    gga = s.func_name.upper().startswith("PBE")
    hbm_peak = peaks.get("hbm_gbs")
    # per-kernel rooflines: algorithmic work per step / CUDA-event time of that kernel.
    # The contractions are charged 2 nbe^2 npts flops each for LDA and GGA alike (the LDA kernels
    # execute about half of that: triangular quadratic form for rho, SYRK for VXC).
    colloc_bytes = st["sum_nbe_npts"] * 8 * (4 if gga else 1) + 32.0 * st["npts"]

    def tensor_entry(name, ms):
        a = 0.5 * f_dense / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        return {"kernel": name, "bound": "tensor", "achieved": a, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": a / dmma_peak, "ms_per_step": ms, "launches_per_step": nb,
                "avg_launch_ms": ms / nb, "flops_per_launch": 0.5 * f_dense / nb}

    per_kernel = [
        {"kernel": "collocation_kernel", "bound": "hbm", "achieved": colloc_bytes / (k_ms[0] * 1e-3) / 1e9
         if k_ms[0] > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
         "frac": (colloc_bytes / (k_ms[0] * 1e-3) / 1e9 / hbm_peak) if k_ms[0] > 0 and hbm_peak else None,
         "ms_per_step": k_ms[0], "launches_per_step": nb, "avg_launch_ms": k_ms[0] / nb,
         "bytes_per_launch": colloc_bytes / nb},
        tensor_entry("fused_xmat_den_zmat_kernel (X = P_sub B on DMMA + rho/grad rho + functional + Z)", k_ms[1]),
        tensor_entry("vxc_kernel (VXC_sub = B^T Z + Z^T B on DMMA + scatter-add)", k_ms[3]),
    ]
    dom = max(per_kernel[1:], key=lambda e: e["ms_per_step"])  # the dominant kernel of the step
    traffic = None
    try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get(args.workload, {}).get(dom["kernel"].split(" ")[0])
    except Exception:
        pass
    dense_ms = k_ms[1] + k_ms[3]
    roofline = {"bound": "tensor", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dmma_peak,
                "unit": "TFLOP/s", "frac": dom["frac"],
                "peak_source": "FP64 DMMA (mma.sync m8n8k4) register-resident probe run in this process; "
                               "MEASURED_PEAKS.json holds HBM/bf16 only (tcgen05 has no FP64 kind)",
                "flops_per_launch": dom["flops_per_launch"], "launches_per_step": nb,
                "avg_launch_ms": dom["avg_launch_ms"], "traffic": traffic,
                "both_contractions_tflops": f_dense / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else 0.0,
                "kernel_ms_per_step": dict(zip(k_names, k_ms)),
                "whole_path_fp64_frac": f_dense / (ms_per_step * 1e-3) / 1e12 / dmma_peak,
                "hbm_peak_gbs": hbm_peak, "hbm_peak_source": peak_src, "per_kernel": per_kernel}

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        tasks = s.lb.export_tasks()
        cpu = oracle_sample(s, tasks, args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "blas", "gflops")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(s, world, "per-step working set (B/dB/Z workspace, %.1f GB streamed) "
                                          "far exceeds the 126 MB L2; no flush needed" %
                                          (st["sum_nbe_npts"] * 8 * (6 if s.func_name.startswith("PBE") else 3) / 1e9)),
                "grid_points": npts_total, "wall_ms_per_step": wall_ms / args.steps,
                "exc": exc_dev, "n_el": nel_dev, "ssf_weights_ms": ssf_ms,
                "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        capi.nccl_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
