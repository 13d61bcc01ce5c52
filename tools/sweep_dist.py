"""Per-gene CD sweep-count distribution from the oracle (diagnostics for the CD kernel's work distribution)."""
import sys; sys.path.insert(0, '.')
import ctypes as C
import numpy as np
from insider_b200 import synth
from oracle import oracle
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pb = synth.ageing_like(N=377, P=P, K=23)
F0, V0 = synth.init_factors(pb.levels, 23, P)
buf = np.zeros((iters + 2, P), dtype=np.int32)
oracle.lib().oracle_set_sweep_sink(buf.ctypes.data_as(C.POINTER(C.c_int)), C.c_longlong(buf.size))
r = oracle.optimize(pb.Y, F0, V0, pb.confounder, None, None, None, 0, 23, 10.0, 10.0, 0.4, 0, 1e-9, 1e-5, iters, perm_mode=1, seed=1)
oracle.lib().oracle_set_sweep_sink(None, C.c_longlong(0))
np.save('/tmp/sweeps.npy', buf)
for it in range(r.iters_run):
    s = buf[it]
    q = np.percentile(s, [0, 10, 50, 90, 99, 100]).astype(int)
    # lockstep efficiency for warps of 32 consecutive genes, and after sorting by previous iteration's count
    def eff(order):
        ss = s[order]; n = len(ss) // 32 * 32
        w = ss[:n].reshape(-1, 32)
        return w.sum() / (w.max(1).sum() * 32)
    e0 = eff(np.arange(P))
    e1 = eff(np.argsort(buf[it - 1], kind='stable')) if it > 0 else float('nan')
    cor = np.corrcoef(buf[it - 1], s)[0, 1] if it > 0 else float('nan')
    print(f"it {it:3d} mean {s.mean():8.1f} pct(0,10,50,90,99,100) {q} lockstep eff natural {e0:.2f} sorted-by-prev {e1:.2f} corr {cor:.2f}", flush=True)
