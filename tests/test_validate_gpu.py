"""Memory / race checking without compute-sanitizer (closed on this pool): scripts/sanitize_case.py runs
every kernel family on small cases with MIPCU_PARAM_VALIDATE = 1 (device-side verification of every
index structure the kernels dereference) and, on two GPUs, repeats the row-sharded solve demanding
bit-identical results across repeats and ranks."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import gen

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_every_kernel_family_on_validated_structures(built):
    p = subprocess.run([sys.executable, str(ROOT / "scripts" / "sanitize_case.py")], capture_output=True, text=True,
                       cwd=ROOT, timeout=900)
    assert p.returncode == 0 and "sanitize_case done" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_validation_passes_on_real_models_and_catches_a_planted_index(built):
    """Not vacuous: MIPCU_PARAM_VALIDATE = 2 plants one out-of-range column index in the device copy before
    validating and the call must fail with the structure named; = 1 passes on every model family."""
    from miplib_b200 import Engine
    from miplib_b200.capi import MipcuError
    for L in (gen.config(1), gen.set_cover(200, 500, 8, 20264), gen.mkp(5, 2000, 4, 10, 20265)):
        with Engine(0) as e:
            e.set_param("validate", 1)
            e.load_lp(L)
            e.prepare()
            v = np.random.default_rng(0).standard_normal(L.n)
            assert np.abs(e.spmv(v) - L.A @ v).max() <= 1e-9 * (1 + np.abs(L.A @ v).max())
        with Engine(0) as e:
            e.set_param("validate", 2)
            e.load_lp(L)
            with pytest.raises(MipcuError, match="index structure violation"):
                e.prepare()


def test_two_rank_protocol_is_deterministic(built):
    from miplib_b200 import Engine
    if Engine.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, str(ROOT / "scripts" / "sanitize_case.py"), "sharded"], capture_output=True,
                       text=True, cwd=ROOT, timeout=1200)
    assert p.returncode == 0 and "sanitize_case done" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("bit-identical") == 2
