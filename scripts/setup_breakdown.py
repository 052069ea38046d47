"""Wall-clock breakdown of the end-to-end path on one config: `python scripts/setup_breakdown.py [cfg]`."""
import sys, time
sys.path.insert(0, ".")
from oracle import gen
from miplib_b200 import Engine
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = gen.config(cfg)
with Engine(0) as e:
    for rep in range(2):
        t0 = time.perf_counter(); e.load_lp(L); t1 = time.perf_counter()
        e.set_param("verbose", 1 if rep else 0)
        e.prepare(); t2 = time.perf_counter()
        e.set_param("verbose", 0)
        s = e.solve(); t3 = time.perf_counter()
        x = e.x(); y = e.y(); t4 = time.perf_counter()
        print(f"rep {rep}: load {1e3*(t1-t0):.1f} ms  prepare {1e3*(t2-t1):.1f} ms  solve {1e3*(t3-t2):.1f} ms (iter {1e3*s.iter_seconds:.1f})  readback {1e3*(t4-t3):.1f} ms", flush=True)
