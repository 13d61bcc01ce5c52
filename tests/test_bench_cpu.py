"""The reference arm of bench.py (CPU only: the restated reference on the host cores) prints one well-formed JSON line with the
contract's keys, on the same `config` as the GPU arm; the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-genes", "96"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "als_iterations_per_second" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["config"]["workload"] == "ageing_full_377x44477_K23_fit" and d["config"]["P"] == 44477 and d["config"]["K"] == 23
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["extrapolated"] is True and "96 of 44477" in cb["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    pc = d["port_vs_reference_sources"]                    # the timed port against the reference's own sources (oracle/_ref), when present
    assert pc is None or pc.get("ok") is True, pc


def test_reference_arm_uses_an_explicit_openmp_team_under_a_launcher():
    """torchrun exports OMP_NUM_THREADS=1: the reference arm must still report (and use) the team it set itself."""
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-genes", "64"], env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


def test_non_zero_ranks_of_the_reference_arm_exit_without_work():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-genes", "64"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "0", "--no-ttt", "--no-late", "--no-tune", "--no-cpu-baseline", "--no-parity"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
