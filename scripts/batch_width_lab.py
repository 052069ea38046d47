"""Lane width of the batched node-LP kernels (miplib_b200/csrc/batch.cu): P = 1, 2, 4 problems per lane and
rows per warp, same 128 node LPs of config 4 / 5, one round each to 1e-6 from the root relaxation's point.
Reports the in-situ kernel times (all nodes in flight), the round time, and checks that every width returns
bit-identical x for every node.   usage: python scripts/batch_width_lab.py [--config 4] [--nodes 128]"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from miplib_b200 import Engine
from oracle import gen

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--nodes", type=int, default=128)
ap.add_argument("--iter-limit", type=int, default=60000)
ap.add_argument("--settings", default="1:8:0,1:8:1,2:0:0,2:0:1,4:0:0,4:0:1",
                help="comma-separated width:rows_per_warp(0 = auto):feed(0 uniform loads, 1 shuffle-fed)")
args = ap.parse_args()

L = gen.config(args.config)
depth = 12 if args.config == 4 else 30
rng = np.random.default_rng(1000)
B = args.nodes
lb = np.tile(L.lb, (B, 1)); ub = np.tile(L.ub, (B, 1))
for b in range(B):
    cols = rng.choice(L.n, depth, replace=False)
    vals = (rng.random(depth) < (0.5 if args.config == 4 else 0.1)).astype(float)
    lb[b, cols] = vals; ub[b, cols] = vals

with Engine(0) as eng:
    eng.load_lp(L)
    eng.set_param("iter_limit", args.iter_limit)
    root = eng.solve()
    xr, yr = eng.x(), eng.y()
    x0 = np.ascontiguousarray(np.tile(xr, (B, 1)))
    y0 = np.ascontiguousarray(np.tile(yr, (B, 1)))
    w0 = np.full(B, root.primal_weight)
    ref_x = None
    for setting in args.settings.split(","):
        width, rpw, feed = setting.split(":")
        os.environ["MIPCU_BATCH_WIDTH"] = width
        os.environ["MIPCU_BATCH_FEED"] = feed
        if int(rpw):
            os.environ["MIPCU_BATCH_RPW"] = rpw
        else:
            os.environ.pop("MIPCU_BATCH_RPW", None)
        t0 = time.perf_counter()
        st, x = eng.solve_batch(lb, ub, want_x=True, x0=x0, y0=y0, weight0=w0)[:2]
        secs = time.perf_counter() - t0
        if ref_x is None:
            ref_x = x.copy()
        rec = {"config": args.config, "nodes": B, "width": int(width), "rows_per_warp": int(rpw) or "auto", "feed": int(feed),
               "row_kernel_us": st[0].row_kernel_us, "col_kernel_us": st[0].col_kernel_us,
               "round_seconds": secs, "node_lps_per_s": B / secs,
               "iterations_mean": float(np.mean([s.iterations for s in st])),
               "iterations_max": int(max(s.iterations for s in st)),
               "optimal": int(sum(s.status == "optimal" for s in st)),
               "bit_identical_to_first_setting": bool(np.array_equal(x, ref_x)),
               "max_abs_x_difference": float(np.abs(x - ref_x).max())}
        print(json.dumps(rec), flush=True)
