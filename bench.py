#!/usr/bin/env python
"""bench.py -- MINCO trajectories optimized / second (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: every rank optimizes its own 65 536
seeded 8-piece corridor problems (BASELINE.json configs[2]; S=3 = MINCO_S3NU, K=16 half-planes
per piece, vel/acc/jerk penalties, fp64) with ONE persistent sm_100a kernel launch
(lbfgs_optimize around costFunctional, allocnet_b200/csrc/lbfgs_device.cuh) and, when there is
more than one rank, one NCCL all-gather of the solved coefficients (configs[4] at 8 ranks).

  value        whole-job trajectories/s, inputs resident in HBM when the timed region starts
  e2e          same metric through the HOST-pointer C-ABI (mincob_set_problems + mincob_optimize
               [_sharded]) from pinned host buffers, H2D/D2H copies inside the timed region
  roofline     algorithmic HBM bytes of the optimize kernel / its CUDA-event duration
  cpu_baseline the CPU oracle (restated MINCO cost functional driven by the reference's own
               gcopter/lbfgs.hpp compiled verbatim, oracle/_ref) on a bounded sample, all host threads
  --impl reference   times that CPU path alone (rank 0), same metric/config

Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MINCO trajectories optimized/sec"
UNIT = "trajectories/s"


def bytes_eval(N: int, K: int, S: int = 3) -> int:
    """Algorithmic HBM bytes of one cost evaluation of one trajectory:
    read x, head, tail, half-planes; write f, g."""
    n = 4 * N - 3
    return 16 * n + 8 + 2 * (3 * S * 8) + 32 * N * K


def bytes_traj_once(N: int, S: int = 3) -> int:
    """Once per optimized trajectory: Trajectory-order coefficient write."""
    return N * 3 * 2 * S * 8


CONFIGS = {  # BASELINE.json configs[1..4] (configs[0] is the CPU-only single solve)
    2: dict(batch=4096, pieces=8, K=0),                       # energy-only
    3: dict(batch=65536, pieces=8, K=16),                     # headline: the configuration the metric is quoted on
    4: dict(batch=65536, pieces=16, K=16, warm_start="net"),  # learned time-allocation warm start (seq5 conv-lstm)
    5: dict(batch=65536, pieces=8, K=16),                     # per GPU; 8 ranks = 524 288 problems + all-gather
}


def workload_name(a) -> str:
    pen = "corridor K=%d + vel/acc/jerk penalties" % a.K if a.K > 0 else "energy-only"
    ws = ("; initial durations from the reference's seq5 conv-lstm net on sliding 5-piece windows (trapezoid rule where "
          "the planner would reject the net's answer)") if a.warm_start == "net" else ""
    fz = "; durations FIXED (MINCOB_FLAG_FREEZE_TIMES)" if a.freeze_times else ""
    return (f"configs[{a.config}]: batch {a.batch} x {a.pieces}-piece per GPU, {pen}, S={a.S} "
            f"({'MINCO_S3NU jerk' if a.S == 3 else 'MINCO_S4NU snap'}), fp64, L-BFGS mem_size={a.mem_size}{ws}{fz}")


def host_info() -> dict:
    model = "?"
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cpu_model": model, "logical_cpus": os.cpu_count()}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.th = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def rd():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.th = threading.Thread(target=rd, daemon=True)
        self.th.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


def make_params(a):
    from allocnet_b200.params import default_params, energy_only
    from allocnet_b200 import params as P
    p = default_params(a.S, mem_size=a.mem_size)
    p.mapping = {"auto": P.MAP_AUTO, "throughput": P.MAP_THROUGHPUT, "latency": P.MAP_LATENCY}[getattr(a, "mapping", "auto")]
    if getattr(a, "freeze_times", False):
        p.flags |= P.FLAG_FREEZE_TIMES
    return p if a.K > 0 else energy_only(p)


def cpu_reference_run(a, sample: int, threads: int | None = None, first: int = 0):
    """The CPU path on `sample` problems of the same workload: restated MINCO cost functional
    under the reference's own L-BFGS (oracle/_ref, gcopter/lbfgs.hpp verbatim) when it was built,
    else the restated driver.  Returns (seconds, result dict, threads, kind, driver)."""
    from allocnet_b200 import synth
    from oracle.pyoracle import Oracle
    orc = Oracle()
    threads = threads or orc.hardware_threads() or (os.cpu_count() or 1)
    pb = synth.make_problems(sample, N=a.pieces, K=a.K, S=a.S, first=first)
    prm = make_params(a)
    t0 = time.perf_counter()
    res = orc.optimize_batch_ref(prm, pb, nthreads=threads)
    dt = time.perf_counter() - t0
    return dt, res, threads, "port", res["driver"]


def run_reference(a):
    """--impl reference: the CPU path alone, K timed steps of a bounded sample, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = a.cpu_sample or 2048
    for w in range(a.warmup):
        cpu_reference_run(a, min(sample, 256), first=1 << 20)
    tot, trajs, evals, thr, kind, drv = 0.0, 0, 0, 1, "port", ""
    for k in range(a.steps):
        dt, res, thr, kind, drv = cpu_reference_run(a, sample, first=k * sample)
        tot += dt; trajs += sample; evals += int(res["evals"].sum())
    v = trajs / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a) + f"; CPU arm: each step is a bounded sample of {sample} problems of that "
                               "workload's seeded stream (trajectories/s does not depend on the batch size on the CPU: "
                               "problems are independent and run one per thread)",
                   "sample_per_step": sample, "host": host_info()},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": kind,
                         "sample": f"{sample} problems per step x {a.steps} steps of the same seeded stream; driver: {drv}; "
                                   "cost functional: oracle/minco_oracle.hpp (restated MINCO, -O3 -march=x86-64-v3)",
                         "evals_per_s": evals / tot, "mean_evals_per_traj": evals / trajs, "host": host_info()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def sass_fingerprint():
    """Identity of the kernel the library was built from: sha256 over the CUDA sources (allocnet_b200/csrc).  The fp64 flop
    count per evaluation in profiles/optimize_kernel_traffic.json is only valid for the kernel it was captured from."""
    import hashlib
    d = os.path.join(ROOT, "allocnet_b200", "csrc")
    h = hashlib.sha256()
    try:
        for name in sorted(os.listdir(d)):
            with open(os.path.join(d, name), "rb") as fh:
                h.update(name.encode()); h.update(fh.read())
        return h.hexdigest()[:16]
    except OSError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS),
                    help="BASELINE.json configs[] preset (3 = headline); --batch/--pieces/--K override it")
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU per step")
    ap.add_argument("--pieces", type=int, default=None)
    ap.add_argument("--K", type=int, default=None)
    ap.add_argument("--S", type=int, default=3)
    ap.add_argument("--mem-size", dest="mem_size", type=int, default=8)
    ap.add_argument("--warm-start", dest="warm_start", default=None, choices=["trapezoid", "net"])
    ap.add_argument("--cpu-sample", dest="cpu_sample", type=int, default=0)
    ap.add_argument("--mapping", default="auto", choices=["auto", "throughput", "latency"],
                    help="mincob_params.mapping: lanes-per-trajectory layout of the optimize kernel")
    ap.add_argument("--freeze-times", action="store_true", help="MINCOB_FLAG_FREEZE_TIMES: durations fixed (the reference's call)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the parity / gather checks (they run outside the timed region)")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the two-batches-in-flight measurement (`pipelined`)")
    a = ap.parse_args()
    preset = CONFIGS[a.config]
    a.batch = a.batch if a.batch is not None else preset["batch"]
    a.pieces = a.pieces if a.pieces is not None else preset["pieces"]
    a.K = a.K if a.K is not None else preset["K"]
    a.warm_start = a.warm_start or preset.get("warm_start", "trapezoid")
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    from allocnet_b200 import api, sharded, synth
    from allocnet_b200 import params as PR

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    prm = make_params(a)
    B, N, K, S = a.batch, a.pieces, a.K, a.S
    n = 4 * N - 3
    # rank r owns problems [r*B, (r+1)*B) of the seeded stream (block partition)
    lo, hi = sharded.shard_range(world * B, world, rank)
    pb = synth.make_problems(hi - lo, N=N, K=K, S=S, first=lo)

    # a real (non-legacy) stream shared by torch and the handle, so torch.cuda.Event brackets the launches
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sh = sharded.ShardedMinco(prm, local, rank, world, dist if world > 1 else None, stream=stream.cuda_stream)
    mb = sh.mb

    # ---- device-resident leg -------------------------------------------------------------
    t = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    d_head, d_tail = t(pb.head), t(pb.tail)
    d_hp = t(pb.hpolys) if K > 0 else None
    d_hr = t(pb.hrows) if K > 0 else None
    d_x0 = t(pb.x0())
    d_x = torch.empty_like(d_x0)
    d_f = torch.empty(B, dtype=torch.float64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    d_iters = torch.empty(B, dtype=torch.int32, device=dev)
    d_evals = torch.empty(B, dtype=torch.int32, device=dev)
    d_T = torch.empty(B, N, dtype=torch.float64, device=dev)
    cnt = B * N * 3 * 2 * S
    d_all = torch.empty(world * cnt, dtype=torch.float64, device=dev)   # gathered coefficients, rank-major
    d_coeffs = d_all[rank * cnt:(rank + 1) * cnt]    # the kernel writes this rank's block in place: in-place ncclAllGather
    mb.set_problems_device(B, N, K, d_head, d_tail, d_hp, d_hr)

    # config 4: the reference's trained time-allocation net gives the initial durations; it runs
    # INSIDE the step, batched on the device, on sliding 5-piece windows
    net = None
    if a.warm_start == "net":
        if K == 0:
            raise SystemExit("--warm-start net needs corridor rows (K > 0)")
        from allocnet_b200 import timealloc
        net = {"w": timealloc.load_weights_npz(device=dev), "q0": t(pb.q0), "T0": t(pb.T0), "ta": timealloc,
               "accepted": None, "ms": []}

    def net_x0():
        T, acc = net["ta"].warm_start_durations_torch(net["w"], d_head, d_tail, d_hp, d_hr, net["q0"], net["T0"])
        net["accepted"] = acc
        big = torch.sqrt(torch.clamp(2.0 * T - 1.0, min=0.0)) - 1.0           # backwardT
        small = 1.0 - torch.sqrt(torch.clamp(2.0 / T - 1.0, min=0.0))
        d_x[:, :N] = torch.where(T > 1.0, big, small)
        d_x[:, N:] = d_x0[:, N:]

    kernel_ms, gather_ev = [], []

    def step(record: bool):
        if net is not None:
            e_n0, e_n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_n0.record(stream)
            net_x0()
            e_n1.record(stream)
            if record:
                net["ms"].append((e_n0, e_n1))
        else:
            d_x.copy_(d_x0)                               # fresh start point (x is in/out)
        mb.optimize_device(d_x, d_f, d_status, d_iters, d_evals, d_coeffs, d_T)
        if world > 1:
            e_g0, e_g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_g0.record(stream)
            mb.allgather_device(d_coeffs, d_all, cnt)     # sendbuf == recvbuf + rank*count: in place
            e_g1.record(stream)
            if record:
                gather_ev.append((e_g0, e_g1))
        if record:
            kernel_ms.append(mb.last_kernel_ms()[0])      # CUDA events around the launch, on its stream

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(a.warmup):
        step(False)
    fence()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step(True)
    e1.record(stream)
    fence()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_tot = float(tms.item())
    mapping = {PR.MAP_THROUGHPUT: "throughput (one lane per piece)", PR.MAP_LATENCY: "latency (one warp per trajectory)"}.get(
        mb.last_mapping() if hasattr(mb.L, "mincob_last_mapping") else 0, "?")
    # attribution of the step on every rank: optimize kernel, all-gather (device events), max / min over ranks
    k_ms_rank = float(np.mean(kernel_ms))
    g_ms_rank = float(np.mean([x.elapsed_time(y) for x, y in gather_ev])) if gather_ev else 0.0
    per_rank = torch.tensor([k_ms_rank, g_ms_rank], dtype=torch.float64, device=dev)
    ranks_ms = [per_rank.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(ranks_ms, per_rank)
    ranks_ms = torch.stack(ranks_ms).cpu().numpy()
    evals_t = d_evals.to(torch.int64).sum().reshape(1)
    if world > 1:
        dist.all_reduce(evals_t)
    evals_all = int(evals_t.item())
    evals_sum = int(d_evals.to(torch.int64).sum().item())
    iters_mean = float(d_iters.to(torch.float64).mean().item())
    status = d_status.cpu().numpy()
    ok_frac = float((status >= 0).mean())
    codes, cnts = np.unique(status, return_counts=True)
    status_hist = {str(int(c)): int(k) for c, k in zip(codes, cnts)}
    evals_np = d_evals.cpu().numpy()

    # ---- checks, outside the timed region -------------------------------------------------------------------
    parity = None
    if not a.no_check:
        parity = {}
        # (1) the device's cost at its final x against the CPU oracle at the same x, first `nchk` problems of this rank
        from oracle.pyoracle import Oracle
        nchk = min(B, 4096)
        orc = Oracle()
        xf = d_x[:nchk].cpu().numpy()
        fo, _ = orc.cost_batch(prm, pb.slice(0, nchk), xf, nthreads=orc.hardware_threads() or 1)
        fd = d_f[:nchk].cpu().numpy()
        okp = status[:nchk] != PR.LBFGSERR_INVALID_FUNCVAL
        rel = np.abs(fd - fo)[okp] / np.abs(fo)[okp]
        mr = torch.tensor([float(rel.max()) if rel.size else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(mr, op=dist.ReduceOp.MAX)
        over = torch.tensor([int((rel > 1e-9).sum())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(over)
        parity.update({"max_rel": float(mr.item()), "n": int(nchk) * world, "tolerance": 1e-9, "n_over_tolerance": int(over.item()),
                       "what": "device cost at the device's final x vs oracle cost_batch at that x (every rank, max)"})
        if S == 4:
            parity["note"] = ("MINCO_S4NU: converged batches contain a few collapsed pieces (T of 0.03 s next to 3 s); there the device's "
                              "junction-state formulation loses digits against the banded LU (DESIGN.md section 1): such problems exceed "
                              "1e-9, the rest hold it")
        # (2) gathered array: every rank checks that each rank's block of ITS gathered copy carries that rank's own
        #     checksum (a checksum of checksums), and that its own block is bit-identical to what its kernel wrote
        if world > 1:
            blocks = d_all.view(torch.int64).reshape(world, cnt)
            sums = blocks.sum(dim=1)                                   # wrapping int64 sums of the bit patterns
            mine = sums[rank].clone().reshape(1)
            want = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(want, mine)
            good = torch.tensor([int(all(int(w.item()) == int(sums[r].item()) for r, w in enumerate(want)))],
                                dtype=torch.int64, device=dev)
            # coefficients of problem p evaluated at t = 0 must be the waypoint the optimizer returned for it
            pos0 = blocks.view(torch.float64).reshape(world, B, N, 3, 2 * S)[rank, :, 1:, :, 2 * S - 1]
            good &= int(torch.equal(pos0, d_x[:, N:].reshape(B, N - 1, 3))) if N > 1 else 1
            dist.all_reduce(good, op=dist.ReduceOp.MIN)
            parity["gather_ok"] = bool(int(good.item()))
            parity["gather_check"] = "int64 checksum of every rank's block in every rank's gathered copy == that rank's own; start positions of the gathered pieces == returned waypoints"
        else:
            parity["gather_ok"] = True
            parity["gather_check"] = "single rank: no collective"

    # ---- quality of what was produced (outside the timed region): sampled limits and corridor residual ----
    d_rep = torch.empty(B, 4, dtype=torch.float64, device=dev)
    d_co = d_coeffs.contiguous()
    mb.check_feasibility_device(d_co, d_T, 32, d_rep)
    torch.cuda.synchronize(dev)
    rep = d_rep.cpu().numpy()
    d_rates = torch.empty(B, 3, dtype=torch.float64, device=dev)
    mb.max_rates_device(d_co, d_T, d_rates)                            # exact maxima (Trajectory::getMaxVelRate ...)
    torch.cuda.synchronize(dev)
    rates = d_rates.cpu().numpy()
    quality = {"samples_per_piece": 33,
               "exact_v_within_2pct": float((rates[:, 0] <= 1.02 * float(prm.v_max)).mean()),
               "exact_a_within_2pct": float((rates[:, 1] <= 1.02 * float(prm.a_max)).mean()),
               "exact_max_speed_p99": float(np.percentile(rates[:, 0], 99)),
               "v_within_2pct": float((rep[:, 0] <= 1.02 * float(prm.v_max)).mean()),
               "a_within_2pct": float((rep[:, 1] <= 1.02 * float(prm.a_max)).mean()),
               "j_within_2pct": float((rep[:, 2] <= 1.02 * float(prm.j_max)).mean()),
               "corridor_within_2cm": float((rep[:, 3] <= 0.02).mean()) if K > 0 else None,
               "max_speed_p99": float(np.percentile(rep[:, 0], 99))}

    # ---- config 4: the same batch from the trapezoid start, for comparison (untimed in `value`) ----------------
    warm = None
    if net is not None:
        acc = net["accepted"]
        net_ms = float(np.mean([x.elapsed_time(y) for x, y in net["ms"]]))
        d_x.copy_(d_x0)
        mb.optimize_device(d_x, d_f, d_status, d_iters, d_evals, d_coeffs, d_T)
        trap_ms = mb.last_kernel_ms()[0]
        warm = {"model": "seq5_tokenthresh0_35 (state_dict exported to tests/golden/timealloc_seq5.npz)",
                "windows_per_trajectory": int(acc.shape[1]), "accepted_window_fraction": float(acc.double().mean().item()),
                "accepted_trajectory_fraction": float(acc.all(dim=1).double().mean().item()),
                "net_forward_ms_per_step": net_ms, "optimize_ms_net_start": k_ms_rank,
                "mean_evals_net_start": evals_sum / B,
                "optimize_ms_trapezoid_start": float(trap_ms),
                "mean_evals_trapezoid_start": float(d_evals.to(torch.float64).mean().item()),
                "note": "windows whose net answer the planner would reject (a duration < 1e-10 on a used segment, i.e. the stop "
                        "token fired early: learning_planner.hpp:181-189) keep the trapezoid rule"}

    # ---- two batches in flight (single rank): a second handle on a second stream works on the next batch while the first
    #      one is in its launch tail.  Not the headline (`value` times strictly sequential steps); it shows what the
    #      device sustains when batches stream in, which is what the tail costs per batch. ---------------------------------
    pipelined = None
    if world == 1 and not a.no_pipeline and net is None:
        from allocnet_b200 import api as _api
        s2 = torch.cuda.Stream(dev)
        mb2 = _api.MincoBatch(prm, device=local)
        mb2.set_stream(s2.cuda_stream)
        mb2.set_problems_device(B, N, K, d_head, d_tail, d_hp, d_hr)
        bufs = []
        for _ in range(2):
            bufs.append(dict(x=torch.empty_like(d_x0), f=torch.empty_like(d_f), st=torch.empty_like(d_status), it=torch.empty_like(d_iters),
                             ev=torch.empty_like(d_evals), co=torch.empty(cnt, dtype=torch.float64, device=dev), T=torch.empty_like(d_T)))
        lanes = ((mb, stream, bufs[0]), (mb2, s2, bufs[1]))

        def pstep(k):
            h, st_, b = lanes[k % 2]
            with torch.cuda.stream(st_):
                b["x"].copy_(d_x0)
                h.optimize_device(b["x"], b["f"], b["st"], b["it"], b["ev"], b["co"], b["T"])
        for k in range(4):
            pstep(k)
        torch.cuda.synchronize(dev)
        nst = max(2 * a.steps, 6)
        p0, p1, pj = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
        p0.record(stream)
        s2.wait_event(p0)
        for k in range(nst):
            pstep(k)
        pj.record(s2)
        stream.wait_event(pj)
        p1.record(stream)
        torch.cuda.synchronize(dev)
        pms = p0.elapsed_time(p1)
        same = bool(torch.equal(bufs[0]["x"], d_x) and torch.equal(bufs[1]["x"], d_x)) if not a.no_check else None
        pipelined = {"value": B * nst / (pms * 1e-3), "unit": UNIT, "steps": nst, "ms_per_step": pms / nst, "batches_in_flight": 2,
                     "results_identical_to_sequential": same,
                     "note": "two handles on two streams, alternating batches: the next batch's blocks become resident as the previous "
                             "batch's blocks retire, so its launch tail is filled; latency of one batch is unchanged"}
        mb2.close()

    # ---- the per-evaluation kernel on its own (the lbfgs_evaluate_t body for the whole batch, one launch): this is the
    #      launch BASELINE.json's "one fused kernel per L-BFGS evaluation" describes, with its algorithmic HBM bytes -------
    evk = None
    if rank == 0:
        d_g = torch.empty_like(d_x0)
        for _ in range(3):
            mb.evaluate_device(d_x0, d_f, d_g)
        torch.cuda.synchronize(dev)
        ev_ms = []
        for _ in range(10):
            mb.evaluate_device(d_x0, d_f, d_g)
            ev_ms.append(mb.last_kernel_ms()[0])          # CUDA events around the launch, on its stream; synchronises
        ev = float(np.median(ev_ms))
        ev_bytes = B * bytes_eval(N, K, S)
        evk = {"kernel": f"evaluate_kernel<S={S}>", "ms_per_launch": ev, "evals_per_s": B / (ev * 1e-3),
               "algorithmic_bytes_per_launch": int(ev_bytes), "achieved_gbs": ev_bytes / (ev * 1e-3) / 1e9}

    # ---- e2e leg: host pointers through the C-ABI, copies inside the timed region -------------
    e2e = None
    if not a.no_e2e:
        pin = api.pinned_empty
        h_head, h_tail = pin(pb.head.shape), pin(pb.tail.shape)
        h_hp = pin(pb.hpolys.shape) if K > 0 else None
        h_hr = pin(pb.hrows.shape, np.int32) if K > 0 else None
        h_x0, h_x = pin((B, n)), pin((B, n))
        h_f, h_T = pin((B,)), pin((B, N))
        h_status, h_iters, h_evals = pin((B,), np.int32), pin((B,), np.int32), pin((B,), np.int32)
        # Where the job's result lands: ONE host array [world*B][N][3][2S] (rank-major) in the consumer's (rank 0's) memory.
        # With several ranks it is a shared-memory segment mapped and page-locked by every rank, and every rank copies its own
        # block device -> host into it over its own PCIe link (mincob_optimize_sharded_local); the device-side all-gather of
        # the step is the same collective as in the device-resident leg.
        shm = None
        if world > 1:
            from multiprocessing import shared_memory
            box = [None]
            if rank == 0:
                shm = shared_memory.SharedMemory(create=True, size=world * cnt * 8)
                box[0] = shm.name
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                shm = shared_memory.SharedMemory(name=box[0])
            h_call_all = np.ndarray((world * cnt,), dtype=np.float64, buffer=shm.buf)
            api.host_register(h_call_all)
            h_call = h_call_all[rank * cnt:(rank + 1) * cnt]
        else:
            h_call_all = h_call = pin((cnt,))
        h_head[...] = pb.head; h_tail[...] = pb.tail; h_x0[...] = pb.x0()
        if K > 0:
            h_hp[...] = pb.hpolys; h_hr[...] = pb.hrows

        class _PB:  # the host arrays in the C-ABI layouts, as LearningPlanner would hand them over
            pass
        hpb = _PB(); hpb.S, hpb.N, hpb.K, hpb.B = S, N, K, B
        hpb.head, hpb.tail, hpb.hpolys, hpb.hrows = h_head, h_tail, h_hp, h_hr

        def e2e_step():
            h_x[...] = h_x0
            mb.set_problems_async(hpb)                                 # H2D: head, tail, hpolys, hrows, chunked, overlapped by the kernel
            # H2D x; optimize; all-gather on the device; D2H of this rank's results into the shared result array
            mb.optimize_sharded_local_host_buffers(h_x, h_f, h_status, h_iters, h_evals, h_call, h_T)
        for _ in range(2):
            e2e_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()                                                 # returns after the results are on the host
        fence()                                                        # every rank's block is in the consumer's array
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ok = True
        if not a.no_check and rank == 0:                               # the consumer sees every rank's block
            got = h_call_all.reshape(world, B, N, 3, 2 * S)
            e2e_ok = bool(np.isfinite(got).all()) and bool((np.abs(got).reshape(world, -1).max(axis=1) > 0).all())
            e2e_ok = e2e_ok and bool(np.array_equal(got[0, :, 1:, :, 2 * S - 1], h_x[:, N:].reshape(B, N - 1, 3)) if N > 1 else True)
        h2d = pb.head.nbytes + pb.tail.nbytes + (pb.hpolys.nbytes + pb.hrows.nbytes if K > 0 else 0) + B * n * 8
        d2h = B * n * 8 + B * 8 + 3 * B * 4 + cnt * 8 + B * N * 8
        e2e = {"value": world * B * a.steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(dt.item()) / a.steps,
               "api": "mincob_set_problems_async + mincob_optimize_sharded_local (host pointers, pinned; the kernel starts while the batch is still uploading)" +
                      ("; bytes are per rank; every rank copies its own block of the result into one shared, page-locked host "
                       "array owned by rank 0 (the consumer), after the device-side all-gather" if world > 1 else ""),
               "ok_fraction": float((h_status >= 0).mean()), "result_complete_on_consumer": e2e_ok}
        if shm is not None:
            fence()
            api.host_unregister(h_call_all)
            del h_call, h_call_all
            shm.close()
            if rank == 0:
                shm.unlink()
        mb.set_problems_device(B, N, K, d_head, d_tail, d_hp, d_hr)

    if rank == 0:
        k_ms = float(ranks_ms[:, 0].max())                 # the slowest rank's kernel is the one the step waits for
        k_ms0 = float(ranks_ms[0, 0])
        alg = evals_sum * bytes_eval(N, K, S) + B * bytes_traj_once(N, S)
        peaks, peak_src = None, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
            peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            peak = 6650.0
        achieved = alg / (k_ms0 * 1e-3) / 1e9
        traffic, fp64 = None, None
        try:  # dram bytes and fp64 flops of one launch from the committed `ncu --set full` capture, if any
            with open(os.path.join(ROOT, "profiles", "optimize_kernel_traffic.json")) as fh:
                tj = json.load(fh)
            if tj.get("batch") == B and tj.get("pieces") == N and tj.get("K") == K and tj.get("S", 3) == S:
                traffic = tj.get("dram_bytes_per_launch")
                stale = tj.get("kernel_source_sha256_16") not in (None, sass_fingerprint())
                if tj.get("fp64_flops_per_eval") and not stale:
                    # what actually bounds the kernel: executed fp64 flops (2 per DFMA, 1 per DADD/DMUL, predicated-on
                    # threads only; counted by ncu per evaluation for this configuration) over the live kernel time,
                    # against the DFMA rate measured on this device a moment ago
                    fl = float(tj["fp64_flops_per_eval"]) * evals_sum
                    pk = mb.measure_fp64_peak()
                    fp64 = {"bound": "fp64", "achieved": fl / (k_ms0 * 1e-3) / 1e12, "peak": pk, "unit": "TFLOP/s",
                            "frac": fl / (k_ms0 * 1e-3) / 1e12 / pk, "flops_per_eval": float(tj["fp64_flops_per_eval"]),
                            "peak_source": "mincob_measure_fp64_peak (independent DFMA chains, CUDA events, this run)",
                            "flops_source": tj.get("source")}
                elif stale:
                    traffic = None
                    print("[bench] profiles/optimize_kernel_traffic.json was captured from another version of allocnet_b200/csrc "
                          "(kernel_source_sha256_16 differs): traffic / roofline_fp64 withheld; re-run tools/update_profiles.py",
                          file=sys.stderr)
        except Exception as e:
            print(f"[bench] no traffic / fp64 record: {e}", file=sys.stderr)
        line = {
            "metric": METRIC, "value": world * B * a.steps / (ms_tot * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_tot / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "batch_per_gpu": B, "pieces": N, "K": K, "S": S,
                       "kappa": int(prm.kappa), "mem_size": int(prm.mem_size), "past": int(prm.past),
                       "delta": float(prm.delta), "mapping": mapping, "warm_start": a.warm_start,
                       "parallelism": f"problems block-partitioned over {world} rank(s)" +
                                      (", one in-place NCCL all-gather of coefficients per step" if world > 1 else ""),
                       "l2": "inputs larger than L2 (half-planes %.0f MB per GPU per step), no flush" % (pb.hpolys.nbytes / 1e6),
                       "host": host_info()},
            "evals_per_s": evals_all * a.steps / (ms_tot * 1e-3),
            "mean_evals_per_traj": evals_sum / B, "p95_evals_per_traj": float(np.percentile(evals_np, 95)),
            "mean_iters_per_traj": iters_mean, "ok_fraction": ok_frac, "status_hist": status_hist, "quality": quality,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": f"optimize_kernel<S={S}>", "kernel_ms": k_ms0,
                         "algorithmic_bytes_per_launch": int(alg), "peak_source": peak_src,
                         "note": "algorithmic bytes = sum(evals)*bytes_eval + B*coeff bytes (SURVEY 8d), rank 0's launch; the persistent "
                                 "kernel keeps a problem on chip for all of its evaluations, so DRAM traffic is ~230x lower; what bounds "
                                 "it is fp64 latency and instruction supply (roofline_fp64, profiles/)"},
            "step_breakdown_ms": {"optimize_kernel_max_over_ranks": k_ms, "optimize_kernel_min_over_ranks": float(ranks_ms[:, 0].min()),
                                  "optimize_kernel_per_rank": [float(v) for v in ranks_ms[:, 0]],
                                  "allgather_max_over_ranks": float(ranks_ms[:, 1].max()),
                                  "allgather_per_rank": [float(v) for v in ranks_ms[:, 1]],
                                  "note": "CUDA events on the launching stream; a rank's all-gather time includes waiting for the slowest "
                                          "rank's kernel (the collective cannot finish before every rank has entered it)"},
            "gpu_launches": a.steps,
            "clocks": clk,
        }
        if parity is not None:
            line["parity_check"] = parity
        if warm is not None:
            line["warm_start"] = warm
        if fp64 is not None:
            line["roofline_fp64"] = fp64
        if evk is not None:
            evk["frac_of_hbm_peak"] = evk["achieved_gbs"] / peak
            line["evaluate_kernel"] = evk
        if pipelined is not None:
            line["pipelined"] = pipelined
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not a.no_cpu:
            sample = a.cpu_sample or 4096
            dt, res, thr, kind, drv = cpu_reference_run(a, sample)
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": thr, "kind": kind,
                                    "sample": f"first {sample} problems of the same seeded stream, {thr} threads; driver: {drv}",
                                    "evals_per_s": float(res["evals"].sum()) / dt,
                                    "mean_evals_per_traj": float(res["evals"].mean()), "host": host_info()}
        print(json.dumps(line), flush=True)
    mb.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
