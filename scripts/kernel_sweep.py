"""Times the hot kernels on one config for each variant: `python scripts/kernel_sweep.py [cfg]`."""
import sys, json
sys.path.insert(0, ".")
from oracle import gen
from miplib_b200 import Engine
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = gen.config(cfg)
out = {}
with Engine(0) as e:
    e.load_lp(L)
    e.prepare()
    for tma in (0,):
        e.set_param("use_tma", tma)
        for tpr in (0, 8):
            e.set_param("threads_per_row", tpr)
            for which, nm in ((2, "row_step"), (3, "col_step")):
                us, nb = e.time_kernel(which, 30, 0)
                usf, _ = e.time_kernel(which, 10, 1)
                out[f"tma{tma}_tpr{tpr}_{nm}"] = (round(us, 1), round(usf, 1), round(nb / us / 1e3))
                print(f"tma{tma} tpr{tpr} {nm}: hot {us:.1f} us, flushed {usf:.1f} us, {nb/us/1e3:.0f} GB/s", flush=True)
    e.set_param("threads_per_row", 0)
    for tma in (0,):
        e.set_param("use_tma", tma)
        s = e.solve()
        print("solve tma", tma, s.status, s.iterations, round(s.solve_seconds, 3), "s row/col us", round(s.row_kernel_us, 1), round(s.col_kernel_us, 1))
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/sweep_c{cfg}.json", "w"), indent=1)
