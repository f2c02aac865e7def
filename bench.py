#!/usr/bin/env python
"""bench.py -- cohort NLML+gradient evaluations per second (BASELINE.json metric).

Workload at every N (weak scaling, per-GPU work fixed): configs[1] of BASELINE.json --
24-feature SM-LMC (D=24, Q=5, R=8, P=1114), 256 synthetic patients of n=500 points per GPU,
one theta per patient drawn from the reference's initialisation distribution.  One "step" is
one NLML+gradient evaluation of every patient of the shard (256 evaluations per GPU).

  value  evaluations/s, theta resident in HBM, CUDA events on the library's stream
  e2e    evaluations/s through medgp_cuda_nlml_grad with HOST buffers (theta H2D and
         nlml/grad/status D2H inside the timed region)
  --impl reference   the reference's own CPU implementation (oracle/_ref/ref_eval, the
         unmodified reference compiled against OpenBLAS) on the box's host cores, one
         single-thread process per core as the reference is deployed.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from medgp_b200 import synth  # noqa: E402

Q, D, R, N_POINTS, PATIENTS_PER_GPU = 5, 24, 8, 500, 256
METRIC = "cohort NLML+gradient evals/sec"
UNIT = "evals/s"


def workload_config(n_gpus):
    return {
        "workload": "C2: 24-feature SM-LMC (feature_all.json shape), Q=5 R=8 P=1114, "
                    f"{PATIENTS_PER_GPU} synthetic patients x n={N_POINTS} per GPU, "
                    "1 NLML+gradient eval per patient per step",
        "patients_per_gpu": PATIENTS_PER_GPU, "n_points": N_POINTS, "Q": Q, "D": D, "R": R,
        "sharding": f"patients sharded over {n_gpus} GPU(s), no collective on the data path",
        "l2": "inputs_exceed_l2 (256 x 2 MiB matrices per step vs 126 MB L2)",
    }


def make_shard(rank):
    """Patients rank*256 .. rank*256+255 of the synthetic cohort and their thetas."""
    patients = [synth.make_patient(D, N_POINTS, seed=rank * PATIENTS_PER_GPU + i)
                for i in range(PATIENTS_PER_GPU)]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, PATIENTS_PER_GPU, seed=718 + rank)
    return patients, thetas


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread
    every ~2 ms (the timed region is tens of milliseconds, too short for `nvidia-smi -lms`)."""

    def __init__(self, index):
        self.index, self.thread, self.stop_flag = index, None, False
        self.sm, self.reasons, self.smax, self.err = [], 0, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as ex:  # noqa: BLE001 -- reported in the JSON line
            self.err = f"nvml unavailable: {ex}"
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception as ex:  # noqa: BLE001
                self.err = str(ex)
                return
            time.sleep(0.002)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not sampled"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        nv = self.nv
        names = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        reasons = sorted(nm for nm, bit in names if self.reasons & int(bit))
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": reasons, "samples": len(self.sm)}


def measure_fp64_peak(torch, device):
    """cuBLAS DGEMM 8192^3 through torch.matmul: the FP64 yardstick (MEASURED_PEAKS.json has
    no FP64 entry).  Returns TFLOP/s (best of 5)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(device)
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def cholesky_metric(api, fp64_peak, n=4000, count=32, reps=3):
    """BASELINE.json secondary metric, 'FP64 % of peak (Cholesky)' on a long-stay series:
    `count` initialisations of one n-point 24-feature series in flight, NLML only; potrf time
    from per-stage CUDA events; flops = n^3/3 per matrix."""
    meta, x, y = synth.make_patient(D, n, seed=4000, T=1200.0)
    thetas = synth.init_hyp_lmc_sm(Q, D, R, count, seed=4)
    ctx = api.Context(Q, D, R, workspace_bytes=12 << 30)
    sid = ctx.add_series(meta, x, y)
    for _ in range(2):
        ctx.nlml_grad([sid] * count, thetas, False)
    ctx.stage_times(reset=True)
    ctx.profile(True)
    for _ in range(reps):
        _, _, st = ctx.nlml_grad([sid] * count, thetas, False)
    t = ctx.stage_times()
    ctx.close()
    ms = t["potrf"]["ms"] / reps
    tf = count * n ** 3 / 3.0 / (ms * 1e-3) / 1e12
    return {"n": n, "matrices_in_flight": count, "potrf_ms": ms, "tflops": tf, "peak_tflops": fp64_peak,
            "frac": tf / fp64_peak, "ok": bool((st == 0).all()),
            "peak_source": "cuBLAS DGEMM 8192^3 measured in this run"}


def reference_sample(n_evals, want_grad=True):
    """Times `n_evals` evaluations of the workload with the compiled reference, one
    single-thread process per evaluation, all concurrently.  Returns (evals/s, cores, text)."""
    from oracle import oracle
    if not oracle.have_ref():
        return None, 0, "oracle/_ref/ref_eval missing"
    patients, thetas = make_shard(0)
    tmp = tempfile.mkdtemp(prefix="medgp_ref_")
    paths = []
    for i in range(n_evals):
        meta, x, y = patients[i % PATIENTS_PER_GPU]
        p = os.path.join(tmp, f"case{i}.txt")
        oracle.write_case(p, Q, D, R, meta, x, y, thetas[i % PATIENTS_PER_GPU])
        paths.append(p)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([oracle.REF_EVAL, p, "1" if want_grad else "0", "1", "1"],
                              stdout=subprocess.DEVNULL, env=oracle.ref_env()) for p in paths]
    for pr in procs:
        pr.wait()
    dt = time.perf_counter() - t0
    for p in paths:
        os.unlink(p)
    os.rmdir(tmp)
    return n_evals / dt, n_evals, (f"{n_evals} concurrent single-thread ref_eval processes, one "
                                  f"NLML+grad eval each (n={N_POINTS}, D={D}, Q={Q}, R={R}), "
                                  f"{dt:.1f} s wall")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        reference_sample(cores)
    t0 = time.perf_counter()
    total = 0
    sample = ""
    for _ in range(args.steps):
        v, used, sample = reference_sample(cores)
        if v is None:
            print(json.dumps({"impl": "reference", "unavailable": sample}))
            return 0
        total += used
    dt = time.perf_counter() - t0
    value = total / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample + "; reference built with g++ -O2 + OpenBLAS (not icpc/MKL)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from medgp_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    patients, thetas = make_shard(rank)
    ctx = api.Context(Q, D, R, device=local, workspace_bytes=8 << 30)
    sids = np.array([ctx.add_series(*p) for p in patients], dtype=np.int32)
    P, B = ctx.P, len(sids)
    d_theta, d_nlml = ctx.malloc(B * P * 8), ctx.malloc(B * 8)
    d_grad, d_status = ctx.malloc(B * P * 8), ctx.malloc(B * 4)
    ctx.h2d(d_theta, thetas)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def step_device():
        ctx.nlml_grad_device(sids, d_theta, True, d_nlml, d_grad, d_status)

    for _ in range(args.warmup):
        step_device()
    ctx.sync()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    ctx.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # ---- the same K steps again with per-stage CUDA events (single stream, stages serialised):
    #      source of the roofline numbers and of the launch count
    ctx.stage_times(reset=True)
    ctx.profile(True)
    ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ep0.record(stream)
    for _ in range(args.steps):
        step_device()
    ep1.record(stream)
    ctx.sync()
    ms_profiled = ep0.elapsed_time(ep1)
    stages = ctx.stage_times(reset=True)
    ctx.profile(False)
    status = np.empty(B, dtype=np.int32)
    ctx.d2h(status, d_status)
    nlml_dev = np.empty(B)
    ctx.d2h(nlml_dev, d_nlml)
    assert (status == 0).all() and np.isfinite(nlml_dev).all(), "device path produced failures"

    # ---- e2e: host buffers through the public C ABI call, copies inside the timed region
    # (theta and the result arrays live in page-locked host memory, as an optimiser loop
    # that calls the backend every iteration would keep them)
    theta_pin = ctx.pinned(thetas.shape)
    theta_pin[...] = thetas
    outs = (ctx.pinned((B,)), ctx.pinned((B, P)), ctx.pinned((B,), np.int32))
    for _ in range(3):
        ctx.nlml_grad(sids, theta_pin, True, out=outs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nlml_h, grad_h, st_h = ctx.nlml_grad(sids, theta_pin, True, out=outs)
    t_e2e = time.perf_counter() - t0
    barrier()
    assert np.allclose(nlml_h, nlml_dev, rtol=1e-12, atol=0)

    t = torch.tensor([ms, t_e2e * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    total_evals = world * B * args.steps
    value = total_evals / (ms_max * 1e-3)
    e2e_value = total_evals / (e2e_ms_max * 1e-3)

    if rank == 0:
        # ---- roofline of the dominant kernel (by device time in the profiled pass)
        names = [k for k in stages if k != "evals"]
        dom = max(names, key=lambda k: stages[k]["ms"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fp64_peak = measure_fp64_peak(torch, dev)
        kernel_of = {"potrf": "k_potrf_panel + k_potrf_diag", "diag": "k_potrf_diag", "trtri": "k_trtri_row", "lauum": "k_lauum",
                     "grad": "k_grad (+k_grad_finish)", "assemble": "k_assemble", "prep": "k_prep",
                     "solve": "k_solve", "predict": "k_cross/k_pred_finish"}
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except OSError:
            pass

        def stage_roofline(name):
            st = dict(stages[name])
            if name == "potrf":  # the diagonal-block kernel belongs to the factorisation
                st["ms"] += stages["diag"]["ms"]
                st["launches"] += stages["diag"]["launches"]
            launches = max(1, st["launches"])
            avg_s = max(st["ms"], 1e-9) * 1e-3 / launches
            if name in ("potrf", "diag", "trtri", "lauum", "solve"):
                a = st["flops"] / launches / avg_s / 1e12
                return {"kernel": kernel_of[name], "bound": "tensor", "achieved": a, "peak": fp64_peak,
                        "unit": "TFLOP/s", "frac": a / fp64_peak}
            a = st["bytes"] / launches / avg_s / 1e9
            return {"kernel": kernel_of[name], "bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s",
                    "frac": a / hbm}

        roofline = stage_roofline(dom)
        roofline["traffic"] = traffic.get(dom)
        roofline["peak_source"] = ("FP64 tensor (DMMA): cuBLAS DGEMM 8192^3 measured in this run; "
                                   "MEASURED_PEAKS.json has no FP64 entry" if roofline["bound"] == "tensor" else hbm_src)
        roofline["algorithmic_work"] = ("SURVEY.md section 8d: potrf n^3/3 flop (panel + diagonal kernels together), trtri/lauum n^3/3 "
                                        "flop each, assembly 8n(n+1)/2+12n B, gradient 8n(n+1)/2+8P B per evaluation, n=500")
        if dom == "grad":
            roofline["fp64_pipe_active_pct_ncu"] = traffic.get("grad_fp64_pipe_pct")
        roofline["measured_in"] = ("second pass of the same K steps with per-stage CUDA events on one "
                                   f"stream ({ms_profiled / args.steps:.3f} ms/step serialised vs "
                                   f"{ms_max / args.steps:.3f} ms/step in the timed multi-stream region)")
        roofline["stage_ms_per_step"] = {k: stages[k]["ms"] / args.steps for k in names}
        roofline["stage_launches_per_step"] = {k: stages[k]["launches"] / args.steps for k in names}
        roofline["all_stages"] = {k: stage_roofline(k) for k in names if stages[k]["ms"] > 0 and k not in ("prep", "solve", "diag")}
        roofline["note"] = ("assembly and gradient kernels are FP64-pipe bound at Q=5 (about 19 FMA per byte against a "
                            "2.6 FMA/B machine balance), so their HBM fraction is low by construction; DESIGN.md section 4")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, used, sample = reference_sample(os.cpu_count() or 1)
            if v is not None:
                cpu = {"value": v, "unit": UNIT, "cores": used, "kind": "reference",
                       "sample": sample + "; reference built with g++ -O2 + OpenBLAS (not icpc/MKL)"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * B * P * 8),
                    "d2h_bytes_per_step": int(world * (B * (P + 1) * 8 + B * 4))},
            "gpu_launches": int(sum(stages[k]["launches"] for k in names)),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        if world == 1 and not args.no_cholesky:
            out["cholesky_fp64"] = cholesky_metric(api, fp64_peak)
        print(json.dumps(out))
    for p in (d_theta, d_nlml, d_grad, d_status):
        ctx.free(p)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cholesky", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
