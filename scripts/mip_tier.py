"""MIP parity tier through the re-hosted reference API (Solver / Var / Constr, Backend::Cuda):
`python scripts/mip_tier.py [names...]` - each model is written as a text fixture and solved by
miplib_b200/miplib_unit_test; the HiGHS-proven optimum is the expected objective."""
import json, os, subprocess, sys, tempfile
from pathlib import Path
sys.path.insert(0, ".")
from oracle import gen
from tests.test_frontend import write_models
from miplib_b200 import build

golden = json.load(open("tests/golden/objectives.json"))
ALL = {
    "setcover_200x500_mip": lambda: gen.set_cover(200, 500, 8, 20264),
    "setcover_400x1000_mip": lambda: gen.set_cover(400, 1000, 8, 20264),
    "mkp_5x60_mip": lambda: gen.mkp(5, 60, 4, 4, 20265),
    "mkp_8x80_mip": lambda: gen.mkp(8, 80, 4, 4, 20265),
    "mkp_10x100_mip": lambda: gen.mkp(10, 100, 4, 5, 20265),
}
names = sys.argv[1:] or list(ALL)
for nm in names:
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "models.txt"
        write_models(path, [(nm, ALL[nm](), golden[nm]["objective"], 1e-9)])
        env = {**os.environ, "MIPLIB_TEST_MODELS": str(path),
               "MIPLIB_TEST_TIME_LIMIT": os.environ.get("MIPLIB_TEST_TIME_LIMIT", "300")}
        p = subprocess.run([str(build.UNIT_TEST), "--filter", "Generated models"], capture_output=True, text=True, env=env)
        lines = [l for l in (p.stdout + p.stderr).splitlines() if nm in l or "failed" in l]
        print("\n".join(lines), flush=True)
