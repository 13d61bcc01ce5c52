"""Developer tool: per-kernel device times of the ALS iteration (CUDA events around every launch).
Usage: python tools/gpu_profile.py [P] [K] [tuning] [iters] [N]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from insider_b200 import _cabi, synth  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 23
    tuning = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 31
    N = int(sys.argv[5]) if len(sys.argv) > 5 else 377
    alpha = float(sys.argv[6]) if len(sys.argv) > 6 else 0.4
    pb = synth.ageing_like(N=N, P=P, K=K)
    tr, te = synth.random_masks(N, P)
    F0, V0 = synth.init_factors(pb.levels, K, P)
    ctx = _cabi.Context(0)
    prob = _cabi.HostProblem(pb.Y, pb.confounder, None, tr if tuning else None, te if tuning else None, 0)
    t0 = time.time()
    res = ctx.upload(prob)
    print(f"upload {time.time() - t0:.3f}s")
    opt = _cabi.default_options()
    opt.lambda1 = opt.lambda2 = 10.0
    opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = alpha, tuning, 1e-9, 1e-5, 100000, 1
    for profile in (False, True):
        ctx.set_profile(profile)
        fac = _cabi.HostFactors(F0, V0, K)
        s = res.begin(fac, opt)
        done = False
        tot = 0.0
        trace = []
        for it in range(iters):
            done, ms = s.step(1)
            tot += ms
            trace.append(ms)
            if done:
                break
        prof = s.profile()
        out = s.end()
        print(f"profile={profile}: {len(trace)} iterations, total {tot:.2f} ms, per-iter first {trace[0]:.3f} ms, last {trace[-1]:.3f} ms, "
              f"median {np.median(trace):.3f} ms; sweeps {out['cd_sweeps']} ({out['cd_sweeps'] / P / len(trace):.1f}/gene-iter); loss {out['loss']:.8g}")
        print("  trace ms:", " ".join(f"{t:.2f}" for t in trace))
        if profile:
            for k, (ms, calls) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
                print(f"  {k:20s} {ms:10.3f} ms  {calls:6d} calls  {1e3 * ms / calls:9.1f} us/call")
    long_iters = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    if long_iters:
        for profile in (False, True):
            ctx.set_profile(profile)
            fac = _cabi.HostFactors(F0, V0, K)
            s = res.begin(fac, opt)
            s.step(long_iters)
            if profile:
                s.profile()   # discard: only the tail is of interest
            p0 = s.profile()
            sw0 = 0
            done, ms = s.step(20)
            p1 = s.profile()
            out = s.end()
            print(f"late phase (iterations {long_iters}..{long_iters + 19}) profile={profile}: {ms / 20:.3f} ms/iter; total sweeps so far {out['cd_sweeps']}; loss {out['loss']:.9g}; checks decay {[c['decay'] for c in out['checks']][-3:]}")
            if profile:
                for k in sorted(p1, key=lambda k: -(p1[k][0] - p0.get(k, (0, 0))[0])):
                    dms = p1[k][0] - p0.get(k, (0, 0))[0]; dc = p1[k][1] - p0.get(k, (0, 0))[1]
                    if dc:
                        print(f"  {k:20s} {dms:10.3f} ms  {dc:6d} calls  {1e3 * dms / dc:9.1f} us/call")


if __name__ == "__main__":
    main()
