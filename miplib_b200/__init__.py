"""miplib_b200 - B200-native LP-relaxation backend for gnosis/miplib (Solver::Backend::Cuda).

The product is native: ``libmiplib_cuda.so`` (hand-written sm_100a kernels behind the C ABI of
``include/mipcu.h``) and ``libmiplib.so`` (the C++17 Solver/Var/Expr/Constr front-end with the
Cuda adapter).  This Python package is only the ctypes binding the tests and ``bench.py`` use to
drive that C ABI; it computes nothing itself and there is no CPU fallback: importing
:mod:`miplib_b200.capi` raises if the shared library has not been built.
"""
from .capi import (Engine, MipcuError, Stats, library_path, load_library, PARAM,  # noqa: F401
                   STATUS_NAMES)
