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


def workload_name(a) -> str:
    pen = "corridor K=%d + vel/acc/jerk penalties" % a.K if a.K > 0 else "energy-only"
    return (f"configs[2]: batch {a.batch} x {a.pieces}-piece per GPU, {pen}, S={a.S} "
            f"({'MINCO_S3NU jerk' if a.S == 3 else 'MINCO_S4NU snap'}), fp64, L-BFGS mem_size={a.mem_size}")


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
    p = default_params(a.S, mem_size=a.mem_size)
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
        "config": {"workload": workload_name(a), "sample_per_step": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": kind,
                         "sample": f"{sample} problems per step x {a.steps} steps of the same seeded stream; driver: {drv}",
                         "evals_per_s": evals / tot, "mean_evals_per_traj": evals / trajs},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--pieces", type=int, default=8)
    ap.add_argument("--K", type=int, default=16)
    ap.add_argument("--S", type=int, default=3)
    ap.add_argument("--mem-size", dest="mem_size", type=int, default=8)
    ap.add_argument("--cpu-sample", dest="cpu_sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    from allocnet_b200 import api, sharded, synth

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
    t = lambda arr: torch.from_numpy(arr).to(dev)
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
    d_coeffs = d_all[rank * cnt:(rank + 1) * cnt] if world == 1 else torch.empty(cnt, dtype=torch.float64, device=dev)
    mb.set_problems_device(B, N, K, d_head, d_tail, d_hp, d_hr)

    kernel_ms = []

    def step(record: bool):
        d_x.copy_(d_x0)                                   # fresh start point (x is in/out)
        sh.optimize_and_gather_device(d_x, d_f, d_status, d_iters, d_evals, d_coeffs, d_T, d_all, cnt)
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
    evals_sum = int(d_evals.to(torch.int64).sum().item())
    iters_mean = float(d_iters.to(torch.float64).mean().item())
    status = d_status.cpu().numpy()
    ok_frac = float((status >= 0).mean())
    codes, cnts = np.unique(status, return_counts=True)
    status_hist = {str(int(c)): int(k) for c, k in zip(codes, cnts)}
    evals_np = d_evals.cpu().numpy()

    # ---- quality of what was produced (outside the timed region): sampled limits and corridor residual ----
    d_rep = torch.empty(B, 4, dtype=torch.float64, device=dev)
    mb.check_feasibility_device(d_coeffs, d_T, 32, d_rep)
    torch.cuda.synchronize(dev)
    rep = d_rep.cpu().numpy()
    d_rates = torch.empty(B, 3, dtype=torch.float64, device=dev)
    mb.max_rates_device(d_coeffs, d_T, d_rates)                        # exact maxima (Trajectory::getMaxVelRate ...)
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
        # rank 0 is the consumer of the whole job: it reads back the coefficients of every rank (rank-major); the other
        # ranks take part in the same all-gather and read back their own shard only
        h_call = pin((world * cnt,)) if rank == 0 else pin((cnt,))
        h_head[...] = pb.head; h_tail[...] = pb.tail; h_x0[...] = pb.x0()
        if K > 0:
            h_hp[...] = pb.hpolys; h_hr[...] = pb.hrows

        class _PB:  # the host arrays in the C-ABI layouts, as LearningPlanner would hand them over
            pass
        hpb = _PB(); hpb.S, hpb.N, hpb.K, hpb.B = S, N, K, B
        hpb.head, hpb.tail, hpb.hpolys, hpb.hrows = h_head, h_tail, h_hp, h_hr

        def e2e_step():
            h_x[...] = h_x0
            mb.set_problems(hpb)                                       # H2D: head, tail, hpolys, hrows
            if rank == 0:
                mb.optimize_sharded_host_buffers(h_x, h_f, h_status, h_iters, h_evals, h_call, h_T)   # H2D x; D2H results
            else:
                mb.optimize_sharded_local_host_buffers(h_x, h_f, h_status, h_iters, h_evals, h_call, h_T)
        for _ in range(2):
            e2e_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()                                                 # returns after the results are on the host
        fence()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = pb.head.nbytes + pb.tail.nbytes + (pb.hpolys.nbytes + pb.hrows.nbytes if K > 0 else 0) + B * n * 8
        d2h = B * n * 8 + B * 8 + 3 * B * 4 + world * cnt * 8 + B * N * 8
        e2e = {"value": world * B * a.steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(dt.item()) / a.steps,
               "api": "mincob_set_problems + mincob_optimize_sharded (host pointers, pinned)" +
                      ("; rank 0 reads back all ranks' coefficients (d2h_bytes_per_step is rank 0's), the other ranks "
                       "(mincob_optimize_sharded_local) their own shard" if world > 1 else ""),
               "ok_fraction": float((h_status >= 0).mean())}
        mb.set_problems_device(B, N, K, d_head, d_tail, d_hp, d_hr)

    if rank == 0:
        k_ms = float(np.mean(kernel_ms))
        alg = evals_sum * bytes_eval(N, K, S) + B * bytes_traj_once(N, S)
        peaks, peak_src = None, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
            peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            peak = 6650.0
        achieved = alg / (k_ms * 1e-3) / 1e9
        traffic, fp64 = None, None
        try:  # dram bytes and fp64 flops of one launch from the committed `ncu --set full` capture, if any
            with open(os.path.join(ROOT, "profiles", "optimize_kernel_traffic.json")) as fh:
                tj = json.load(fh)
            if tj.get("batch") == B and tj.get("pieces") == N and tj.get("K") == K and tj.get("S", 3) == S:
                traffic = tj.get("dram_bytes_per_launch")
                if tj.get("fp64_flops_per_eval"):
                    # what actually bounds the kernel: executed fp64 flops (2 per DFMA, 1 per DADD/DMUL, predicated-on
                    # threads only; counted by ncu per evaluation for this configuration) over the live kernel time,
                    # against the DFMA rate measured on this device a moment ago
                    fl = float(tj["fp64_flops_per_eval"]) * evals_sum
                    pk = mb.measure_fp64_peak()
                    fp64 = {"bound": "fp64", "achieved": fl / (k_ms * 1e-3) / 1e12, "peak": pk, "unit": "TFLOP/s",
                            "frac": fl / (k_ms * 1e-3) / 1e12 / pk, "flops_per_eval": float(tj["fp64_flops_per_eval"]),
                            "peak_source": "mincob_measure_fp64_peak (independent DFMA chains, CUDA events, this run)",
                            "flops_source": tj.get("source")}
        except Exception as e:
            print(f"[bench] no traffic / fp64 record: {e}", file=sys.stderr)
        line = {
            "metric": METRIC, "value": world * B * a.steps / (ms_tot * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_tot / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "batch_per_gpu": B, "pieces": N, "K": K, "S": S,
                       "kappa": int(prm.kappa), "mem_size": int(prm.mem_size), "past": int(prm.past),
                       "delta": float(prm.delta),
                       "parallelism": f"problems block-partitioned over {world} rank(s)" +
                                      (", one NCCL all-gather of coefficients per step" if world > 1 else ""),
                       "l2": "inputs larger than L2 (half-planes %.0f MB per GPU per step), no flush" % (pb.hpolys.nbytes / 1e6)},
            "evals_per_s": world * evals_sum * a.steps / (ms_tot * 1e-3),
            "mean_evals_per_traj": evals_sum / B, "p95_evals_per_traj": float(np.percentile(evals_np, 95)),
            "mean_iters_per_traj": iters_mean, "ok_fraction": ok_frac, "status_hist": status_hist, "quality": quality,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": f"optimize_kernel<S={S}>", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": int(alg), "peak_source": peak_src,
                         "note": "algorithmic bytes = sum(evals)*bytes_eval + B*coeff bytes (SURVEY 8d); the persistent kernel keeps "
                                 "a problem on chip for its ~420 evaluations, so DRAM traffic is ~230x lower; what bounds it is "
                                 "fp64 latency and instruction supply (roofline_fp64, profiles/)"},
            "gpu_launches": a.steps,
            "clocks": clk,
        }
        if fp64 is not None:
            line["roofline_fp64"] = fp64
        if evk is not None:
            evk["frac_of_hbm_peak"] = evk["achieved_gbs"] / peak
            line["evaluate_kernel"] = evk
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not a.no_cpu:
            sample = a.cpu_sample or 4096
            dt, res, thr, kind, drv = cpu_reference_run(a, sample)
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": thr, "kind": kind,
                                    "sample": f"first {sample} problems of the same seeded stream, {thr} threads; driver: {drv}",
                                    "evals_per_s": float(res["evals"].sum()) / dt,
                                    "mean_evals_per_traj": float(res["evals"].mean())}
        print(json.dumps(line), flush=True)
    mb.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
