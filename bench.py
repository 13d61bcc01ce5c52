#!/usr/bin/env python
"""Benchmark of the INSIDER alternating-optimisation fit on B200 (BASELINE.json metric: ALS iterations/s and
time-to-global_tol on the ageing-shaped 377 x 44477, K = 23 fit, next to the CPU path).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   # the restated reference (oracle port) on the host cores
  python bench.py --workload ageing_full_377x44477_K23_tune --gpus N    # tune() grid (51 fits) as replicas over N GPUs

A "step" is one ALS iteration (src/optimize.cpp:325-410). The timed region is iterations 0..K-1 of the fit from the
standard N(0, 0.001^2) initialisation (the per-iteration cost of this algorithm depends on the iteration index: the
elastic-net solver needs thousands of sweeps per gene in the first iterations); the W warm-up steps run the same
kernels on the same data in a throw-away session first. Genes are sharded across ranks (strong scaling: the matrix is
fixed); the only exchange is the all-reduce of the row-side sufficient statistics.

Untimed-by-the-driver extra legs of the default run (each reported in the JSON line): per-kernel device times of the same
K iterations, end to end from pinned and from pageable host memory, steady-state iterations, time-to-global_tol, a small
multi-rank fit compared with the CPU oracle (parity_check), the CPU baseline on a bounded sample (N = 1 only).
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
    # BASELINE.json configs[2]: the tune() grid on the same shape (tuning = 1), run as replicas
    "ageing_full_377x44477_K23_tune": dict(N=377, P=44477, K=23, lam=10.0, alpha=0.4, tuning=1),
    "ageing_full_377x44477_K23_masked": dict(N=377, P=44477, K=23, lam=10.0, alpha=0.4, tuning=1),
    "ageing_toy_377x5000_K23_fit": dict(N=377, P=5000, K=23, lam=10.0, alpha=0.4, tuning=0),
}
CPU_SAMPLE_GENES = 4096
NCU_CD_DENSE_DRAM_BYTES = 26169856 + 2304     # profiles/r02_ncu_k_cd_dense_pform_dense_A.txt
FP64_PEAK = 37.1   # TFLOP/s, FP64 pipe peak measured on this pool by tools/microbench.cu (profiles/r01_microbench_fp64_hbm.txt)
CD_FLOPS_PER_UPDATE = lambda K: 2.0 * K + 12.0   # noqa: E731  (q update 2K + scalar chain 12)


def bytes_per_iter(N, P, K, tuning):
    """algorithmic bytes per ALS iteration (two passes over Y, masks at 1 bit, V 3x, U 2x)."""
    if tuning == 1:
        return 2 * (8 * N * P + N * P / 8) + N * P / 8 + 3 * 8 * K * P + 2 * 8 * N * K
    return 2 * 8 * N * P + 3 * 8 * K * P + 2 * 8 * N * K


def config_for(args, w):
    """The workload description, IDENTICAL in both arms (ours / --impl reference)."""
    return {"workload": args.workload, **{k: w[k] for k in ("N", "P", "K", "lam", "alpha", "tuning")},
            "timed_iterations": f"0..{args.steps - 1} from N(0,0.001^2) init", "parallelism": "gene-sharded",
            "l2": "no flush: Y (134 MB) exceeds the 126 MB L2 at N=1; at N>1 the shard is L2-resident in the real fit too"}


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
            time.sleep(0.005)

    def mark(self):
        """Start of the timed region: forget what was sampled before (the thread is started early so that NVML's slow first calls are over)."""
        self.samples, self.reasons = [], set()

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


def cpu_oracle(n_threads=None):
    """The oracle for a TIMED CPU leg: -march=native build compiled on this host, explicit OpenMP team (torchrun exports
    OMP_NUM_THREADS=1). Returns (module, team size really in effect, build kind)."""
    from oracle import oracle
    native = oracle.use_native()
    team = oracle.set_threads(n_threads or (os.cpu_count() or 1))
    return oracle, team, ("-O3 -march=native" if native else "-O3 -march=x86-64-v3 (native build failed)")


def time_cpu(oracle, team, w, P, steps):
    pb, tr, te, F0, V0 = make_problem(w, P=P)
    r = oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, te, 0, w["K"], w["lam"], w["lam"], w["alpha"], w["tuning"], 1e-12, 1e-5,
                        steps - 1, perm_mode=1, seed=1, n_cores_row=team, n_cores_col=team)
    return r


def run_reference(args, w, rank):
    """The restated reference (oracle/insider_oracle.cpp, OpenMP) on the host cores: the same K iterations from the same
    initialisation on the FULL gene set - measured, not extrapolated."""
    if rank != 0:
        return
    oracle, team, flags = cpu_oracle()
    if args.warmup > 0:     # warm-up: thread team start-up, page faults of the allocator (small sample; the fit itself has no cache to warm)
        time_cpu(oracle, team, w, 256, min(args.warmup, 2))
    P = args.ref_genes or w["P"]
    t0 = time.perf_counter()
    r = time_cpu(oracle, team, w, P, args.steps)
    wall = time.perf_counter() - t0
    secs = r.seconds_in_loop
    value = args.steps / secs * (P / w["P"])
    sample = (f"all {w['P']} genes, iterations 0..{args.steps - 1} (measured, no extrapolation)" if P == w["P"] else
              f"first {P} of {w['P']} genes, iterations 0..{args.steps - 1}; EXTRAPOLATED by {P}/{w['P']}")
    # the timed implementation is the PORT (oracle/insider_oracle.cpp); where oracle/_ref exists (the reference's own sources compiled against
    # oracle/ref_shim/, serial) a small fit shows that the port computes what the reference's code computes
    port_check = None
    try:
        from oracle import ref
        from insider_b200 import synth
        if ref.available():
            pbs = synth.ageing_like(N=80, P=120, K=8, n_donors=13, seed=3)
            trs, tes = synth.random_masks(80, 120, 0.1, 6)
            F0s, V0s = synth.init_factors(pbs.levels, 8, 120, seed=7)
            oracle.set_threads(1)
            ro = oracle.optimize(pbs.Y, F0s, V0s, pbs.confounder, None, trs, tes, 0, 8, 3.0, 3.0, 0.4, 1, 1e-12, 1e-5, 11, perm_mode=0, r_seed=42)
            Fr, Vr, _, _, lr = ref.optimize(pbs.Y, F0s, V0s, pbs.confounder, None, trs, tes, 0, 8, 3.0, 3.0, 0.4, 1, 1e-12, 1e-5, 11, r_seed=42)
            oracle.set_threads(team)
            dv = float(np.abs(Vr - ro.column_factor).max() / np.abs(ro.column_factor).max())
            port_check = {"shape": "80x120 K=8, 12 masked elastic-net iterations, same R stream", "max_rel_diff_V": dv,
                          "rel_diff_loss": abs(lr - ro.loss) / abs(ro.loss), "ok": bool(dv < 1e-10)}
    except Exception as e:  # noqa: BLE001
        port_check = {"error": str(e)[:200]}
    line = {"impl": "reference", "metric": "als_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_for(args, w),
            "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": team, "host_cpus": os.cpu_count(), "kind": "port", "sample": sample,
                             "extrapolated": P != w["P"], "sample_seconds": secs, "wall_seconds_incl_setup": wall, "build": flags,
                             "cd_sweeps_per_gene_iter": r.cd_sweeps / P / args.steps, "loss_after_timed": r.loss},
            "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "port_vs_reference_sources": port_check,
            "why_the_port_is_timed": "oracle/_ref (the reference's own sources) is built serial, because the reference draws randperm from one global RNG "
                                     "inside its OpenMP loops, and on an API shim whose products / views copy where Armadillo + BLAS do not: timing it would "
                                     "understate the reference. The OpenMP port is the faster, i.e. the conservative, CPU arm; port_vs_reference_sources "
                                     "shows that it computes what the reference's code computes."}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing (rendezvous, barrier, max over ranks); the data path's only collective is the library's own."""

    def __init__(self, rank, world, local):
        import torch
        self.torch, self.rank, self.world, self.local = torch, rank, world, local
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; libinsider_b200 has no CPU fallback")
        torch.cuda.set_device(local)
        if world > 1:
            import torch.distributed as tdist
            tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
            self.td = tdist

    def barrier(self):
        if self.world > 1:
            self.td.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX if op == "max" else self.td.ReduceOp.SUM)
        return float(t.item())

    def gather_array(self, a):
        """all ranks' float64 arrays of identical shape, stacked on a new leading axis"""
        if self.world == 1:
            return a[None]
        t = self.torch.from_numpy(np.ascontiguousarray(a)).cuda()
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.td.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def close(self):
        if self.world > 1:
            self.td.destroy_process_group()


def parity_check(ctx, dist):
    """A small gene-sharded fit on ALL ranks against the CPU oracle on rank 0 (both tunings): carries multi-GPU parity
    evidence in the bench line (the driver's test box has one GPU). Untimed."""
    from insider_b200 import _cabi, synth
    N, P, K, iters = 96, 1024, 10, 10
    pb = synth.ageing_like(N=N, P=P, K=K, n_donors=17, seed=5)
    tr, te = synth.random_masks(N, P, 0.1, 6)
    F0, V0 = synth.init_factors(pb.levels, K, P, seed=7)
    out = {"shape": f"{N}x{P} K={K}, {iters + 1} iterations, lambda=3 alpha=0.4, world={dist.world}"}
    for tuning in (0, 1):
        prob = _cabi.HostProblem(pb.Y, pb.confounder, None, tr, te, 0)
        fac = _cabi.HostFactors(F0, V0, K)
        opt = _cabi.default_options()
        opt.lambda1 = opt.lambda2 = 3.0
        opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, tuning, 1e-9, 1e-5, iters, 9
        rg = ctx.optimize(prob, fac, opt)
        sweeps = dist.reduce(float(rg["cd_sweeps"]), "sum")
        if dist.rank == 0:
            from oracle import oracle
            oracle.set_threads(min(8, os.cpu_count() or 1))
            ro = oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, te, 0, K, 3.0, 3.0, 0.4, tuning, 1e-9, 1e-5, iters, perm_mode=1, seed=9)
            dv = float(np.abs(fac.V - ro.column_factor).max() / np.abs(ro.column_factor).max())
            da = float(max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(fac.factors, ro.factors)))
            dl = float(abs(rg["loss"] - ro.loss) / abs(ro.loss))
            out[f"tuning{tuning}"] = {"dV": dv, "dA": da, "dloss": dl, "iters_equal": rg["iters_run"] == ro.iters_run,
                                      "sweeps_equal": int(sweeps) == ro.cd_sweeps, "ok": bool(dv < 1e-8 and da < 1e-8 and dl < 1e-10)}
    return out


def run_ours(args, w, rank, world, local):
    from insider_b200 import _cabi, dist as ibdist

    dist = Dist(rank, world, local)
    torch = dist.torch
    ctx = ibdist.make_context(local)
    pb, tr, te, F0, V0 = make_problem(w)
    N, P, K = w["N"], w["P"], w["K"]

    def opts(max_iter, gtol=1e-12):
        o = _cabi.default_options()
        o.lambda1 = o.lambda2 = w["lam"]
        o.alpha, o.tuning, o.global_tol, o.sub_tol, o.max_iter, o.seed = w["alpha"], w["tuning"], gtol, 1e-5, max_iter, 1
        return o

    # pinned host copy of Y for the end-to-end leg
    Yp = torch.empty((P, N), dtype=torch.float64, pin_memory=True)       # row-major (P, N) == column-major (N, P)
    Yp.numpy()[...] = pb.Y.T
    Yhost = Yp.numpy().T
    prob = _cabi.HostProblem(Yhost, pb.confounder, None, tr, te, 0)
    res = ctx.upload(prob)

    # ---- warm-up: same kernels, same data, throw-away session
    sampler = ClockSampler(local)
    sampler.start()
    if args.warmup > 0:
        s = res.begin(_cabi.HostFactors(F0, V0, K), opts(10 ** 6))
        s.step(args.warmup)
        s.end(read_factors=False)

    # ---- timed region: iterations 0..K-1, device-timed inside the library (CUDA events on its stream)
    fac = _cabi.HostFactors(F0, V0, K)
    s = res.begin(fac, opts(10 ** 6))
    dist.barrier()
    sampler.mark()
    t_wall = time.perf_counter()
    done, ms = s.step(args.steps)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    dist.barrier()
    sampler.stop_flag = True
    out = s.end(read_factors=False)
    ms = dist.reduce(ms)
    value = args.steps / (ms * 1e-3)

    # ---- per-kernel device times of the SAME iterations 0..K-1 (separate profiled session: plain launches, one event pair per kernel)
    ctx.set_profile(True)
    sp = res.begin(_cabi.HostFactors(F0, V0, K), opts(10 ** 6))
    n_prof = args.steps
    sp.step(n_prof)
    prof = sp.profile()
    outp = sp.end(read_factors=False)
    ctx.set_profile(False)

    # ---- end to end through the one-shot C-ABI call with host buffers (H2D of Y/masks/factors and D2H of factors inside).
    # One untimed one-shot call of a single iteration first: the call allocates its device buffers and instantiates its CUDA
    # graphs anew every time, and the first such call in a process also grows the driver's memory pools.
    ctx.optimize(prob, _cabi.HostFactors(F0, V0, K), opts(0))
    dist.barrier()
    t0 = time.perf_counter()
    oe = ctx.optimize(prob, _cabi.HostFactors(F0, V0, K), opts(args.steps - 1))
    t_e2e = dist.reduce(time.perf_counter() - t0)
    # the same from PAGEABLE host memory (what R hands over)
    prob_pg = _cabi.HostProblem(pb.Y, pb.confounder, None, tr, te, 0)
    ctx.optimize(prob_pg, _cabi.HostFactors(F0, V0, K), opts(0))     # untimed, like the pinned leg's (first use allocates the library's bounce buffers)
    dist.barrier()
    t0 = time.perf_counter()
    ctx.optimize(prob_pg, _cabi.HostFactors(F0, V0, K), opts(args.steps - 1))
    t_e2e_pg = dist.reduce(time.perf_counter() - t0)

    # ---- steady-state iterations (late phase of the fit: a few sweeps per gene); every rank takes part
    ms_late = None
    if not args.no_late:
        sl = res.begin(_cabi.HostFactors(F0, V0, K), opts(10 ** 6))
        sl.step(args.late_start)
        dist.barrier()
        _, ms_late = sl.step(20)
        sl.end(read_factors=False)
        ms_late = dist.reduce(ms_late)

    # ---- time to global_tol (the other half of BASELINE.json's metric): whole resident fit, wall clock between barriers
    ttt = []
    if not args.no_ttt:
        for gtol, max_iter in ((1e-7, 50000), (1e-9, 50000)):
            dist.barrier()
            t0 = time.perf_counter()
            o = res.optimize(_cabi.HostFactors(F0, V0, K), opts(max_iter, gtol))
            torch.cuda.synchronize()
            secs = dist.reduce(time.perf_counter() - t0)
            ttt.append({"global_tol": gtol, "max_iter": max_iter, "seconds": secs, "iters_run": o["iters_run"],
                        "converged": bool(o["iters_run"] <= max_iter), "device_loop_seconds": o["loop_ms"] * 1e-3,
                        "final_loss": o["loss"], "ms_per_iteration": 1e3 * secs / max(1, o["iters_run"])})

    pc = parity_check(ctx, dist) if not args.no_parity else None

    # ---- BASELINE.json config 3 beside it: the tune() grid (51 masked fits of 31 iterations) as replicas, one context per GPU
    tg = None
    if not args.no_tune:
        tl = tune_leg(args, WORKLOADS["ageing_full_377x44477_K23_tune"], dist, local)
        if rank == 0:
            tg = {k: tl[k] for k in ("fits", "grid_seconds", "fits_per_second", "chosen_rank", "device_loop_seconds_sum_over_fits", "reg_tuning_best")}
            tg["what"] = tl["config"]["grid"] + "; " + tl["config"]["parallelism"] + "; wall clock of the whole grid between barriers"

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
            kern[name] = {"ms_total": kms, "calls": calls, "us_per_call": 1e3 * kms / max(1, calls), "ms_per_step": kms / n_prof}
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
        # dominant kernel: the elastic-net solver. Algorithmic FP64 flops per coordinate update: 2K for the q update + 12 for
        # the soft-threshold / division / loss-decrement chain.
        cd = None
        cd_name = next((n for n in ("k_cd_dense", "k_cd_masked", "k_col_solve") if n in kern), None)
        if cd_name:
            steps_cd, sweeps = outp["cd_steps"], outp["cd_sweeps"]
            secs = kern[cd_name]["ms_total"] * 1e-3
            flops = steps_cd * CD_FLOPS_PER_UPDATE(K)
            cd = {"kernel": cd_name, "gene_sweeps": sweeps, "coordinate_updates": steps_cd, "gene_sweeps_per_s": sweeps / secs,
                  "algorithmic_flops": flops, "fp64_tflops": flops / secs / 1e12, "fp64_peak_tflops_measured": FP64_PEAK,
                  "share_of_iteration": kern[cd_name]["share"], "ms_per_step": kern[cd_name]["ms_per_step"]}
        roof_dom = None
        if cd:
            roof_dom = {"bound": "fp64", "limiter": "FP64 work on the CUDA cores (DFMA chain, no tensor-core form), held below the FP64 peak by the shared-memory data "
                                                    "pipe: l1tex 89 % busy with the broadcast loads of the XtX row, FP64 pipe 40 % (ncu, profiles/"
                                                    "r02_ncu_k_cd_dense_pform_dense_A.txt), and by the lone-warp tail of the longest genes; DESIGN.md 4",
                        "kernel": cd_name, "achieved": cd["fp64_tflops"], "peak": FP64_PEAK,
                        "unit": "TFLOP/s", "frac": cd["fp64_tflops"] / FP64_PEAK,
                        # dram__bytes_read + write of ONE launch from the committed ncu --set full capture (not measured in this run)
                        "traffic": NCU_CD_DENSE_DRAM_BYTES if (cd_name == "k_cd_dense" and world == 1 and P == 44477) else None,
                        "traffic_source": "profiles/r02_ncu_k_cd_dense_pform_dense_A.txt (ALS iteration 3's launch: 26.17 MB read = Xty + V + order + 22 MB of pre-permuted tables, 2 KB written; algorithmic input 18.6 MB)",
                        "peak_source": "FP64 pipe peak measured with DMMA m8n8k4 by tools/microbench.cu on this pool's B200 (MEASURED_PEAKS.json has "
                                       "no FP64 entry; its bf16 tensor peak does not apply to an FP64 path)",
                        "algorithmic_flops_per_launch": flops / max(1, kern[cd_name]["calls"]),
                        "avg_launch_ms": kern[cd_name]["ms_total"] / max(1, kern[cd_name]["calls"]),
                        "launches": kern[cd_name]["calls"], "kernel_ms_per_step": kern[cd_name]["ms_per_step"],
                        "profiled_iterations": f"0..{n_prof - 1} (the timed region's)"}
        roof_iter = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_iteration": b_iter,
                     "what": "whole ALS iteration, SURVEY.md 8(d) bytes / mean iteration time (per GPU)", "dominant_kernel": dominant}
        late = None
        if ms_late is not None:
            late = {"iterations": f"{args.late_start}..{args.late_start + 19}", "ms_per_iteration": ms_late / 20,
                    "hbm_GBps_algorithmic": b_iter / (ms_late / 20 * 1e-3) / 1e9 / world, "frac_of_hbm_peak": b_iter / (ms_late / 20 * 1e-3) / 1e9 / world / hbm_peak}
        line = {
            "metric": "als_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_for(args, w),
            "e2e": {"value": args.steps / t_e2e, "unit": "iterations/s", "h2d_bytes_per_step": oe["h2d_bytes"] / args.steps,
                    "d2h_bytes_per_step": oe["d2h_bytes"] / args.steps, "seconds": t_e2e, "device_loop_seconds": oe["loop_ms"] * 1e-3,
                    "what": "insider_b200_optimize (one-shot C ABI) from pinned host Y; one untimed 1-iteration call of the same entry point first"},
            "e2e_pageable": {"value": args.steps / t_e2e_pg, "unit": "iterations/s", "seconds": t_e2e_pg,
                             "what": "the same call from pageable host memory (what R passes)"},
            "gpu_launches": int(out["kernel_launches"]),
            "clocks": sampler.summary(),
            "roofline": roof_dom or roof_iter,
            "roofline_iteration_hbm": roof_iter,
            "roofline_kernels": {"streaming": stream, "coordinate_descent": cd, "kernels": kern, "profiled_iterations": n_prof},
            "late_phase": late,
            "time_to_global_tol": ttt,
            "parity_check": pc,
            "tune_grid": tg,
            "cd_sweeps_per_gene_iter": out["cd_sweeps"] / max(1, Pl) / args.steps,
            "loss_after_timed": out["loss"], "wall_s_timed_region": t_wall,
        }
        # CPU baseline beside it (rank 0, N = 1 only): the restated reference on a bounded sample of the same workload
        if world == 1 and not args.no_cpu_baseline:
            oracle, team, flags = cpu_oracle()
            Ps = min(CPU_SAMPLE_GENES, P)
            r = time_cpu(oracle, team, w, Ps, args.steps)
            v = args.steps / r.seconds_in_loop * (Ps / P)
            line["cpu_baseline"] = {"value": v, "unit": "iterations/s", "cores": team, "host_cpus": os.cpu_count(), "kind": "port", "build": flags,
                                    "sample": f"first {Ps} of {P} genes, iterations 0..{args.steps - 1}; scaled by {Ps}/{P} (bounded sample; "
                                              "`bench.py --impl reference` measures the full gene set)",
                                    "extrapolated": True, "sample_seconds": r.seconds_in_loop}
        print(json.dumps(line), flush=True)
    res.release()
    ctx.close()
    dist.close()


# ----------------------------------------------------------------------------------------------------------------------
def tune_grid(w):
    """BASELINE.json configs[2] (README.md:79): latent_dimension 10..30 step 2 at (0.1, 0.1, 0), then lambda 1..19 step 2 x
    alpha {0.2, 0.3, 0.4, 0.5} at the chosen rank - 11 + 40 fits of tuning_iter = 30 (31 ALS iterations each)."""
    ranks = list(range(10, 31, 2))
    lams = [float(v) for v in range(1, 20, 2)]
    alphas = [0.2, 0.3, 0.4, 0.5]
    return ranks, lams, alphas


def tune_leg(args, w, dist, local):
    """tune() grid as REPLICAS: every rank holds the full problem, the points of each phase are dealt round-robin in order of
    decreasing expected cost; no data-path communication (the RMSE table is gathered with torch.distributed between phases).
    Returns the JSON line (rank 0) or None."""
    from insider_b200 import _cabi, synth

    rank, world = dist.rank, dist.world
    N, P = w["N"], w["P"]
    pb = synth.ageing_like(N=N, P=P, K=w["K"])
    tr, te = synth.random_masks(N, P, 0.1, 7)
    n_rep = max(1, args.replicas_per_gpu)
    ctxs = [_cabi.Context(local) for _ in range(n_rep)]
    res0 = ctxs[0].upload(_cabi.HostProblem(pb.Y, pb.confounder, None, tr, te, 0))
    residents = [res0] + [_cabi.Resident.share(c, res0) for c in ctxs[1:]]
    ranks, lams, alphas = tune_grid(w)
    tuning_iter = args.tune_iters

    def run_phase(phase, points):
        """points: (K, l1, l2, alpha); this rank runs points[rank::world]; returns the table [n, 4] of (train, test) RMSE, ms, iterations."""
        mine = list(range(rank, len(points), world))
        facs, optl = [], []
        for i in mine:
            K, l1, l2, a = points[i]
            F0, V0 = synth.init_factors(pb.levels, K, P, seed=1000 * phase + i)
            facs.append(_cabi.HostFactors(F0, V0, K))
            o = _cabi.default_options()
            o.lambda1, o.lambda2, o.alpha, o.tuning, o.global_tol, o.sub_tol, o.max_iter, o.seed = l1, l2, a, 1, 1e-9, 1e-5, tuning_iter, 1
            optl.append(o)
        outs, _ = _cabi.tune_batch(residents, facs, optl) if mine else ([], [])
        tab = np.full((len(points), 4), 0.0)
        for i, o in zip(mine, outs):
            tab[i] = (o["train_rmse"], o["test_rmse"], o["loop_ms"], o["iters_run"])
        return dist.gather_array(tab).sum(axis=0)

    # untimed warm-up: one short fit per distinct kernel family (ridge, elastic net)
    if args.warmup > 0:
        for K, a in ((ranks[0], 0.0), (w["K"], 0.4)):
            F0, V0 = synth.init_factors(pb.levels, K, P, seed=3)
            o = _cabi.default_options()
            o.lambda1 = o.lambda2 = 1.0
            o.alpha, o.tuning, o.max_iter = a, 1, args.warmup
            res0.optimize(_cabi.HostFactors(F0, V0, K), o)
    sampler = ClockSampler(local)
    dist.barrier()
    sampler.start()
    t0 = time.perf_counter()
    p1 = [(k, 0.1, 0.1, 0.0) for k in sorted(ranks, reverse=True)]                    # R/insider.R:120-121
    tab1 = run_phase(0, p1)
    best = p1[int(np.argmin(tab1[:, 1]))][0]                                          # which.min(test_rmse)  :136
    p2 = [(best, l, l, a) for a in alphas for l in lams]                              # expand.grid: lambda fastest  :145
    tab2 = run_phase(1, p2)
    dist.barrier()
    secs = dist.reduce(time.perf_counter() - t0)
    sampler.stop_flag = True
    line = None
    if rank == 0:
        n_fits = len(p1) + len(p2)
        iters = float(tab1[:, 3].sum() + tab2[:, 3].sum())
        line = {"metric": "als_iterations_per_second", "value": iters / secs, "unit": "iterations/s", "n_gpus": world, "steps": int(iters),
                "warmup": args.warmup, "ms_per_step": 1e3 * secs / iters, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "ageing_full_377x44477_K23_tune", "N": N, "P": P, "grid": "K 10..30 step 2 @ (0.1,0.1,0), then lambda 1..19 step 2 x alpha {.2,.3,.4,.5}",
                           "tuning": 1, "tuning_iter": tuning_iter, "parallelism": f"replicas: grid points round-robin over ranks, {n_rep} context(s) per GPU",
                           "l2": "no flush: Y (134 MB) exceeds L2"},
                "grid_seconds": secs, "fits": n_fits, "fits_per_second": n_fits / secs, "chosen_rank": int(best),
                "device_loop_seconds_sum_over_fits": float((tab1[:, 2].sum() + tab2[:, 2].sum()) * 1e-3),
                "rank_tuning": [[p[0], float(t[0]), float(t[1])] for p, t in zip(p1, tab1)],
                "fit_device_ms": {"rank_phase_ridge": [round(float(t[2]), 1) for t in tab1], "reg_phase_elastic_net": [round(float(t[2]), 1) for t in tab2]},
                "reg_tuning_best": [p2[int(np.argmin(tab2[:, 1]))][1], p2[int(np.argmin(tab2[:, 1]))][3], float(tab2[:, 1].min())],
                "e2e": {"value": iters / secs, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "what": "wall clock of the whole grid between barriers: factor H2D/D2H of every fit inside, the one upload of Y/masks outside"},
                "gpu_launches": None, "clocks": sampler.summary()}
    for r in residents[1:]:
        r.release()
    res0.release()
    for c in ctxs:
        c.close()
    return line


def run_tune(args, w, rank, world, local):
    dist = Dist(rank, world, local)
    line = tune_leg(args, w, dist, local)
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ageing_full_377x44477_K23_fit", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-late", action="store_true", help="skip the steady-state (late iterations) measurement")
    ap.add_argument("--no-ttt", action="store_true", help="skip the time-to-global_tol legs")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity leg")
    ap.add_argument("--no-tune", action="store_true", help="skip the tune() grid leg (BASELINE.json config 3 as replicas)")
    ap.add_argument("--late-start", type=int, default=150)
    ap.add_argument("--ref-genes", type=int, default=0, help="--impl reference: time only the first G genes (default: all)")
    ap.add_argument("--replicas-per-gpu", type=int, default=2, help="tune workload: concurrent contexts per GPU (2: one fit's tail overlaps the other's bulk, measured 12.3 -> 11.2 s on one GPU)")
    ap.add_argument("--tune-iters", type=int, default=30, help="tune workload: tuning_iter")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {args.gpus}", file=sys.stderr)
    if args.workload.endswith("_tune"):
        run_tune(args, w, rank, world, local)
    else:
        run_ours(args, w, rank, world, local)


if __name__ == "__main__":
    main()
