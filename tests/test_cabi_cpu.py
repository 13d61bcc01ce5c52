"""CPU-side checks of the boundary: the shared library loads and exports every symbol include/insider_b200.h declares,
refuses to compute without a B200 (no CPU fallback), and the product path never touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "insider_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(insider_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from insider_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/insider_b200.h but not exported"
    assert sorted(_cabi.EXPORTED) == syms
    assert lib.insider_b200_version() == 200


def test_struct_layouts_match_header():
    from insider_b200 import _cabi
    o = _cabi.default_options()
    assert (o.lambda1, o.lambda2, o.alpha, o.tuning) == (1.0, 1.0, 0.1, 1)          # src/optimize.cpp:257 defaults
    assert (o.global_tol, o.sub_tol, o.max_iter, o.check_every) == (1e-10, 1e-5, 10000, 10)
    assert ctypes.sizeof(_cabi.Check) == 8 + 9 * 8
    assert ctypes.sizeof(_cabi.Options) == 72
    assert ctypes.sizeof(_cabi.Problem) == 16 + 16 + 5 * 8


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from insider_b200 import _cabi
    with pytest.raises(_cabi.InsiderError) as e:
        _cabi.Context(0)
    assert e.value.code == _cabi.ERR_CUDA
    with pytest.raises(_cabi.InsiderError):
        from insider_b200 import api
        api.set_default_context(None)
        api.optimize(np.zeros((4, 4)), [np.zeros((2, 2))], np.zeros((2, 4)), np.ones((4, 1), np.int32), None, None, None, 0, 2, tuning=0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "insider_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt, f
                assert "/root/reference" not in txt, f


def test_interaction_column_first_appearance_order():
    """R/insider.R:34-40: interaction levels are numbered in first-appearance order and inserted as column 2."""
    from insider_b200 import synth
    a = np.array([2, 1, 2, 1, 1, 2]); b = np.array([1, 1, 2, 1, 2, 1])
    assert synth.interaction_column(a, b).tolist() == [1, 2, 3, 2, 4, 1]


def test_synth_shapes_and_levels():
    from insider_b200 import synth
    p = synth.ageing_like(N=377, P=64, K=23)
    assert p.Y.shape == (377, 64) and p.Y.flags.f_contiguous
    assert p.levels[0] == 2 and p.levels[2] == 8 and p.levels[3] == 107 and p.levels[1] <= 16
    for c, L in enumerate(p.levels):
        assert sorted(np.unique(p.confounder[:, c]).tolist()) == list(range(1, L + 1))      # levels exactly 1..L_c
    q = synth.with_continuous(N=60, P=20, K=4, levels=(3, 4), Q=2)
    assert q.X.shape == (60, 2)


def test_rcpp_shim_type_checks_against_the_c_abi():
    """r-pkg/src/optimize_b200.cpp cannot be built here (no R / Rcpp), but its calls into include/insider_b200.h can be
    type-checked against a stub Rcpp.h: a signature drift between the shim and the C ABI fails this test."""
    import subprocess
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "r-pkg", "src", "optimize_b200.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "r-pkg", "src", "optimize_b200.cpp")).read()
    # the five registered entry points of the reference (src/RcppExports.cpp:112-116) are all defined
    for fn in ("optimize", "strong_coordinate_descent", "coordinate_descent", "optimize_continuous_v2", "optimize_continuous"):
        assert re.search(r"\[\[Rcpp::export\]\]\s*\n[^\n]*\b" + fn + r"\(", src), fn


def test_r_package_configure_finds_the_library(tmp_path):
    import shutil
    import subprocess
    pkg = tmp_path / "pkg"
    shutil.copytree(os.path.join(ROOT, "r-pkg"), pkg)
    env = dict(os.environ, INSIDER_B200_HOME=ROOT)
    r = subprocess.run(["sh", "./configure"], cwd=pkg, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mk = open(pkg / "src" / "Makevars").read()
    assert f"-I{ROOT}/include" in mk and "-linsider_b200" in mk
    env = {k: v for k, v in os.environ.items() if not k.startswith("INSIDER_B200")}
    env["INSIDER_B200_HOME"] = str(tmp_path / "nowhere")
    r = subprocess.run(["sh", "./configure"], cwd=pkg, env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "not found" in r.stderr
