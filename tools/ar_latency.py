"""Developer tool: stream latency of the per-iteration all-reduce ([SB | G], ~69 KB of FP64 at 377 x 44477 K = 23) on N GPUs:
200 back-to-back ncclAllReduce of that size, CUDA events, max over ranks.
Usage: torchrun --nproc-per-node N tools/ar_latency.py [doubles]"""
import os
import sys

import torch
import torch.distributed as td

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8640
x = torch.ones(n, dtype=torch.float64, device="cuda")
for _ in range(20):
    td.all_reduce(x)
torch.cuda.synchronize()


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    td.barrier()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def eager():
    for _ in range(200):
        td.all_reduce(x)


ms_eager = timed(eager, 200)
if td.get_rank() == 0:
    print(f"all-reduce of {n} doubles ({n * 8 / 1024:.1f} KB) on {td.get_world_size()} GPUs: {1e3 * ms_eager:.1f} us each, back to back on one stream")
td.destroy_process_group()
