"""Short driver for ncu: load one LP config, prepare it, launch each hot kernel a few times.
`python scripts/prof_kernels.py [config=3] [reps=3] [tpr=0]`."""
import sys
sys.path.insert(0, ".")
from oracle import gen
from miplib_b200 import Engine

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tpr = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = gen.config(cfg)
with Engine(0) as e:
    e.load_lp(L)
    e.prepare()
    if tpr:
        e.set_param("threads_per_row", tpr)
    for which in (2, 3):
        us, nb = e.time_kernel(which, reps, 0)
        print(which, round(us, 1), "us", round(nb / us / 1e3, 1), "GB/s")
