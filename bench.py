#!/usr/bin/env python
"""Benchmark of the INSIDER alternating-optimisation fit on B200 (BASELINE.json metric: ALS iterations/s on the
ageing-shaped 377 x 44477, K = 23 fit, next to the CPU path).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   # the restated reference (oracle port) on the host cores

A "step" is one ALS iteration (src/optimize.cpp:325-410). The timed region is iterations 0..K-1 of the fit from the
standard N(0, 0.001^2) initialisation (the per-iteration cost of this algorithm depends on the iteration index: the
elastic-net solver needs hundreds of sweeps per gene in the first iterations); the W warm-up steps run the same
kernels on the same data in a throw-away session first. Genes are sharded across ranks (strong scaling: the matrix is
fixed); the only exchange is the all-reduce of the row-side sufficient statistics.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: full ageing-shaped synthetic, fit() => tuning = 0 (R/insider.R:190 partition = 0)
    "ageing_full_377x44477_K23_fit": dict(N=377, P=44477, K=23, lam=10.0, alpha=0.4, tuning=0),
    "ageing_full_377x44477_K23_tune": dict(N=377, P=44477, K=23, lam=10.0, alpha=0.4, tuning=1),
    "ageing_toy_377x5000_K23_fit": dict(N=377, P=5000, K=23, lam=10.0, alpha=0.4, tuning=0),
}
CPU_SAMPLE_GENES = 2048


def bytes_per_iter(N, P, K, tuning):
    """algorithmic bytes per ALS iteration (two passes over Y, masks at 1 bit, V 3x, U 2x)."""
    if tuning == 1:
        return 2 * (8 * N * P + N * P / 8) + N * P / 8 + 3 * 8 * K * P + 2 * 8 * N * K
    return 2 * 8 * N * P + 3 * 8 * K * P + 2 * 8 * N * K


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def make_problem(w, P=None):
    from insider_b200 import synth
    P = P or w["P"]
    pb = synth.ageing_like(N=w["N"], P=P, K=w["K"])
    tr = te = None
    if w["tuning"] == 1:
        tr, te = synth.random_masks(w["N"], P, 0.1, 7)
    F0, V0 = synth.init_factors(pb.levels, w["K"], P, seed=1)
    return pb, tr, te, F0, V0


def run_reference(args, w, rank):
    """The restated reference (oracle/insider_oracle.cpp, OpenMP) on the host cores: K iterations from the same
    initialisation on a bounded sample (the first CPU_SAMPLE_GENES genes); value scaled to the full gene count."""
    if rank != 0:
        return
    from oracle import oracle
    cores = os.cpu_count() or 1
    Ps = min(CPU_SAMPLE_GENES, w["P"])
    pb, tr, te, F0, V0 = make_problem(w, P=Ps)
    if args.warmup > 0:
        oracle.optimize(pb.Y[:, :64], F0, V0[:, :64], pb.confounder, None, None if tr is None else tr[:, :64], None if te is None else te[:, :64],
                        0, w["K"], w["lam"], w["lam"], w["alpha"], w["tuning"], 1e-12, 1e-5, 0, perm_mode=1, seed=1, n_cores_row=cores, n_cores_col=cores)
    r = oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, te, 0, w["K"], w["lam"], w["lam"], w["alpha"], w["tuning"], 1e-12, 1e-5,
                        args.steps - 1, perm_mode=1, seed=1, n_cores_row=cores, n_cores_col=cores)
    secs = r.seconds_in_loop
    value = args.steps / secs * (Ps / w["P"])
    sample = f"first {Ps} of {w['P']} genes, iterations 0..{args.steps - 1}; iterations/s scaled by {Ps}/{w['P']} (cost is linear in genes)"
    line = {"impl": "reference", "metric": "als_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, **{k: w[k] for k in ("N", "P", "K", "lam", "alpha", "tuning")}},
            "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample,
                             "sample_seconds": secs, "cd_sweeps_per_gene_iter": r.cd_sweeps / Ps / args.steps},
            "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, w, rank, world, local):
    import torch
    from insider_b200 import _cabi, dist as ibdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libinsider_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ibdist.make_context(local)
    pb, tr, te, F0, V0 = make_problem(w)
    N, P, K = w["N"], w["P"], w["K"]

    def opts(max_iter):
        o = _cabi.default_options()
        o.lambda1 = o.lambda2 = w["lam"]
        o.alpha, o.tuning, o.global_tol, o.sub_tol, o.max_iter, o.seed = w["alpha"], w["tuning"], 1e-12, 1e-5, max_iter, 1
        return o

    def barrier():
        if world > 1:
            import torch.distributed as tdist
            tdist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as tdist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    # pinned host copy of Y for the end-to-end leg
    Yp = torch.empty((P, N), dtype=torch.float64, pin_memory=True)       # row-major (P, N) == column-major (N, P)
    Yp.numpy()[...] = pb.Y.T
    Yhost = Yp.numpy().T
    prob = _cabi.HostProblem(Yhost, pb.confounder, None, tr, te, 0)
    res = ctx.upload(prob)

    # ---- warm-up: same kernels, same data, throw-away session
    if args.warmup > 0:
        s = res.begin(_cabi.HostFactors(F0, V0, K), opts(10 ** 6))
        s.step(args.warmup)
        s.end(read_factors=False)

    # ---- timed region: iterations 0..K-1, device-timed inside the library (CUDA events on its stream)
    sampler = ClockSampler(local)
    fac = _cabi.HostFactors(F0, V0, K)
    s = res.begin(fac, opts(10 ** 6))
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    done, ms = s.step(args.steps)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    barrier()
    sampler.stop_flag = True
    out = s.end(read_factors=False)
    ms = max_over_ranks(ms)
    value = args.steps / (ms * 1e-3)

    # ---- per-kernel device times (separate short profiled session; not part of the timed number)
    ctx.set_profile(True)
    sp = res.begin(_cabi.HostFactors(F0, V0, K), opts(10 ** 6))
    n_prof = min(args.steps, 12)
    sp.step(n_prof)
    prof = sp.profile()
    outp = sp.end(read_factors=False)
    ctx.set_profile(False)

    # ---- end to end through the one-shot C-ABI call with host buffers (H2D of Y/masks/factors and D2H of factors inside)
    # one untimed one-shot call of a single iteration first: the call allocates its device buffers and instantiates its CUDA
    # graphs anew every time, and the first such call in a process also grows the driver's memory pools (measured: 0.11 s
    # against 0.24 s for the same 25 iterations)
    ctx.optimize(prob, _cabi.HostFactors(F0, V0, K), opts(0))
    barrier()
    fac_e = _cabi.HostFactors(F0, V0, K)
    t0 = time.perf_counter()
    oe = ctx.optimize(prob, fac_e, opts(args.steps - 1))
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = args.steps / t_e2e

    # ---- steady-state iterations (late phase of the fit: a few sweeps per gene); every rank takes part
    ms_late = None
    if not args.no_late:
        sl = res.begin(_cabi.HostFactors(F0, V0, K), opts(10 ** 6))
        sl.step(args.late_start)
        barrier()
        _, ms_late = sl.step(20)
        sl.end(read_factors=False)
        ms_late = max_over_ranks(ms_late)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        b_iter = bytes_per_iter(N, P, K, w["tuning"])
        achieved = b_iter / (ms * 1e-3 / args.steps) / 1e9 / world
        Pl = ibdist.gene_block(P, world, 0)[1]
        kern = {}
        for name, (kms, calls) in prof.items():
            kern[name] = {"ms_total": kms, "calls": calls, "us_per_call": 1e3 * kms / max(1, calls)}
        tot_prof = sum(v["ms_total"] for v in kern.values()) or 1.0
        for v in kern.values():
            v["share"] = v["ms_total"] / tot_prof
        dominant = max(kern, key=lambda k: kern[k]["ms_total"]) if kern else None
        stream = {}
        y_bytes = 8.0 * N * Pl + (N * Pl / 8 if w["tuning"] else 0) + 8.0 * K * Pl
        for kname in ("k_row_b", "k_col_xty", "k_sse"):
            if kname in kern:
                stream[kname] = {"algorithmic_bytes": y_bytes, "GBps": y_bytes / (kern[kname]["us_per_call"] * 1e-6) / 1e9,
                                 "frac_of_hbm_peak": y_bytes / (kern[kname]["us_per_call"] * 1e-6) / 1e9 / hbm_peak}
        # dominant kernel: the persistent elastic-net solver. Algorithmic FP64 flops per coordinate update: 2K for the
        # q -= delta * XtX[:,k] update + 12 for the soft-threshold / exact division / loss-decrement chain.
        FP64_PEAK = 37.1   # TFLOP/s, measured on this pool by tools/microbench.cu (profiles/r01_microbench_fp64_hbm.txt)
        cd = None
        cd_name = "k_cd_dense" if "k_cd_dense" in kern else ("k_col_solve" if "k_col_solve" in kern else None)
        if cd_name:
            steps_cd, sweeps = outp["cd_steps"], outp["cd_sweeps"]
            secs = kern[cd_name]["ms_total"] * 1e-3
            flops = steps_cd * (2.0 * K + 12.0)
            cd = {"kernel": cd_name, "gene_sweeps": sweeps, "coordinate_updates": steps_cd, "gene_sweeps_per_s": sweeps / secs,
                  "algorithmic_flops": flops, "fp64_tflops": flops / secs / 1e12, "fp64_peak_tflops_measured": FP64_PEAK,
                  "share_of_iteration": kern[cd_name]["share"]}
        roof_dom = None
        if cd:
            note = ("thread-per-gene coordinate descent, K=23: every step needs the 24-double table row in every thread; ncu shows the "
                    "shared-memory data pipe at 92 % of peak (l1tex__throughput) with the FP64 pipe at 42 %; lockstep warps run until their "
                    "slowest gene converges (lane efficiency ~0.8 with phased re-grouping / ordering by the previous counts). "
                    "See profiles/r01_ncu_k_cd_dense_v5_dense_A.txt, r01_cd_phases.txt"
                    if cd_name == "k_cd_dense" else
                    "8-lanes-per-gene coordinate descent with per-gene Gram matrices: bound by the shared-memory pipe; "
                    "see profiles/r01_ncu_k_cd_persistent_*.txt")
            # dram__bytes_read.sum + dram__bytes_write.sum of one whole-problem launch from the ncu --set full capture of this
            # command's shape (profiles/r01_ncu_k_cd_dense_v5_dense_A.txt: 20.37 MB read, 0 written; algorithmic Xty + V = 18.6 MB;
            # the phase launches of iterations 0-2 read their share of it)
            traffic = 20372224 if (cd_name == "k_cd_dense" and world == 1 and args.workload == "ageing_full_377x44477_K23_fit") else None
            roof_dom = {"bound": "tensor", "kernel": cd_name, "achieved": cd["fp64_tflops"], "peak": FP64_PEAK,
                        "unit": "TFLOP/s", "frac": cd["fp64_tflops"] / FP64_PEAK, "traffic": traffic,
                        "peak_source": "FP64 pipe peak measured with DMMA m8n8k4 by tools/microbench.cu on this pool's B200 (MEASURED_PEAKS.json has "
                                       "no FP64 entry; its bf16 tensor peak does not apply to an FP64 path)",
                        "note": note,
                        "algorithmic_flops_per_launch": flops / max(1, kern[cd_name]["calls"]),
                        "avg_launch_ms": kern[cd_name]["ms_total"] / max(1, kern[cd_name]["calls"])}
        roof_iter = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_iteration": b_iter,
                     "what": "whole ALS iteration, SURVEY.md 8(d) bytes / mean iteration time (per GPU)", "dominant_kernel": dominant}
        # steady-state iterations (late phase of the fit: a few sweeps per gene), for context
        late = None
        if ms_late is not None:
            late = {"iterations": f"{args.late_start}..{args.late_start + 19}", "ms_per_iteration": ms_late / 20,
                    "hbm_GBps_algorithmic": b_iter / (ms_late / 20 * 1e-3) / 1e9 / world, "frac_of_hbm_peak": b_iter / (ms_late / 20 * 1e-3) / 1e9 / world / hbm_peak}
        line = {
            "metric": "als_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, **{k: w[k] for k in ("N", "P", "K", "lam", "alpha", "tuning")},
                       "timed_iterations": f"0..{args.steps - 1} from N(0,0.001^2) init", "parallelism": f"gene-sharded x{world}",
                       "l2": "Y (134 MB) exceeds L2 at N=1; at N>1 the shard is L2-resident in the real fit too (no flush between iterations)"},
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": oe["h2d_bytes"] / args.steps,
                    "d2h_bytes_per_step": oe["d2h_bytes"] / args.steps, "seconds": t_e2e, "device_loop_seconds": oe["loop_ms"] * 1e-3, "what": "insider_b200_optimize (one-shot C ABI) from pinned host Y; one untimed 1-iteration call of the same entry point first"},
            "gpu_launches": int(out["kernel_launches"]),
            "clocks": sampler.summary(),
            "roofline": roof_dom or roof_iter,
            "roofline_iteration_hbm": roof_iter,
            "roofline_kernels": {"streaming": stream, "coordinate_descent": cd, "kernels": kern, "profiled_iterations": n_prof},
            "late_phase": late,
            "cd_sweeps_per_gene_iter": out["cd_sweeps"] / max(1, Pl) / args.steps,
            "loss_after_timed": out["loss"], "wall_s_timed_region": t_wall,
        }
        # CPU baseline beside it (rank 0, N = 1 only): the restated reference on a bounded sample
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle
            cores = os.cpu_count() or 1
            Ps = min(CPU_SAMPLE_GENES, P)
            pbs, trs, tes, F0s, V0s = make_problem(w, P=Ps)
            r = oracle.optimize(pbs.Y, F0s, V0s, pbs.confounder, None, trs, tes, 0, K, w["lam"], w["lam"], w["alpha"], w["tuning"], 1e-12, 1e-5,
                                args.steps - 1, perm_mode=1, seed=1, n_cores_row=cores, n_cores_col=cores)
            v = args.steps / r.seconds_in_loop * (Ps / P)
            line["cpu_baseline"] = {"value": v, "unit": "iterations/s", "cores": cores, "kind": "port",
                                    "sample": f"first {Ps} of {P} genes, iterations 0..{args.steps - 1}, scaled by {Ps}/{P}",
                                    "sample_seconds": r.seconds_in_loop}
        print(json.dumps(line), flush=True)
    res.release()
    ctx.close()
    if world > 1:
        import torch.distributed as tdist
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ageing_full_377x44477_K23_fit", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-late", action="store_true", help="skip the steady-state (late iterations) measurement")
    ap.add_argument("--late-start", type=int, default=150)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {args.gpus}", file=sys.stderr)
    run_ours(args, w, rank, world, local)


if __name__ == "__main__":
    main()
