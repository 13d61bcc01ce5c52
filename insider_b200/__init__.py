"""insider_b200 — B200-native (sm_100a) implementation of INSIDER's alternating-optimisation fit behind the reference's API.

Package contents: csrc/ (CUDA kernels + C ABI), _cabi.py (ctypes binding), api.py (mirror of the reference's R API),
synth.py (synthetic data of the reference's named shapes).
"""
from . import _cabi, synth  # noqa: F401
from .api import InsiderObject, fit, init_parameters, insider, optimize, ratio_splitter, tune  # noqa: F401

__all__ = ["insider", "tune", "fit", "optimize", "ratio_splitter", "init_parameters", "InsiderObject", "synth"]
