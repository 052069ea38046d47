"""One block (64 iterations) of the batched node-LP engine on config 4 or 5, 128 nodes - short enough for
`ncu --set full` on the batched kernels:  python scripts/batch_profile.py [4|5]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import gen
from miplib_b200 import Engine

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = gen.config(cfg)
rng = np.random.default_rng(5)
B = 128
lb = np.tile(L.lb, (B, 1)); ub = np.tile(L.ub, (B, 1))
for b in range(B):
    cols = rng.choice(L.n, 12, replace=False)
    vals = (rng.random(12) < 0.5).astype(float)
    lb[b, cols] = vals; ub[b, cols] = vals
with Engine(0) as e:
    e.load_lp(L)
    e.set_param("iter_limit", 64)
    st, _ = e.solve_batch(lb, ub, want_x=False)
    print("config", cfg, "row us", st[0].row_kernel_us, "col us", st[0].col_kernel_us, "bytes", st[0].row_kernel_bytes, st[0].col_kernel_bytes)
