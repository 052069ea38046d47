"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mipcu.h declares, and fails loudly (never falls back) when no GPU is present.
No compute call is made here.  (-m "not gpu")"""
import ctypes
import subprocess

import pytest


def test_library_exports_every_declared_symbol(built):
    from miplib_b200 import capi
    lib = capi.load_library()
    declared = capi.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mipcu.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.library_path())],
                         capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported


def test_library_is_sm100a_only(built):
    from miplib_b200 import capi
    out = subprocess.run(["cuobjdump", "--list-elf", str(capi.library_path())],
                         capture_output=True, text=True).stdout
    archs = {tok for ln in out.splitlines() for tok in ln.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_hot_kernels_have_no_spills_and_keep_their_register_budget(built):
    """Resource usage of the shipping kernels from the sm_100a cubin (cuobjdump -res-usage): the two fused
    iteration kernels stay at 40 registers (6 CTAs of 256 threads per SM), the wide-lane batched kernels within
    128 (2 CTAs per SM), none of them with a stack frame - a spill in these loops would show up as LDL / STL."""
    import re
    from miplib_b200 import capi
    out = subprocess.run(["cuobjdump", "-res-usage", str(capi.library_path())], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in ln:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(\w+):(\d+)", ln)}
            cur = None
    budget = {"k_row_stepILi0ELb0E": 40, "k_col_stepILi0ELb0ELb0E": 40, "k_row_step_batch_fedILi4E": 128,
              "k_row_step_batch_fedILi2E": 128, "k_col_step_batchILb0ELb0ELi2E": 128, "k_col_step_batchILb0ELb1ELi2E": 128,
              "k_row_batch_longILb0ELi0ELi2E": 128, "k_row_step_ownILi0ELb0ELb0E": 64, "k_col_step_ownILi0ELb0ELb0ELb0E": 64}
    for pattern, regs in budget.items():
        hits = {k: v for k, v in usage.items() if pattern in k}
        assert hits, pattern
        for name, u in hits.items():
            assert u["REG"] <= regs, (name, u)
            assert u["STACK"] == 0 and u["LOCAL"] == 0, (name, u)


def test_stats_struct_layout_matches_header(built):
    from miplib_b200 import capi
    # 2 x int32 + 2 x int64 + 12 x double + 2 x int64
    assert ctypes.sizeof(capi._Stats) == 8 + 16 + 12 * 8 + 16 + 16   # + col_partial_us, col_exchange_us


def test_no_cpu_fallback_without_gpu(built):
    from miplib_b200 import capi
    if capi.Engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.MipcuError, match="no CPU fallback"):
        capi.Engine(0)


def test_null_context_calls_return_error(built):
    from miplib_b200 import capi
    lib = capi.load_library()
    assert lib.mipcu_solve_lp(None, None) == 1
    assert lib.mipcu_set_param(None, 0, 1.0) == 1
    lib.mipcu_destroy(None)   # must be a no-op
