"""One GPU: programmatic dependent launch on / off for the two iteration kernels.
  gpurun -- python scripts/pdl_lab.py
Config 3 (kernel-bound) and config 2 with direct launches (launch-bound without the CUDA graph)."""
import sys
sys.path.insert(0, ".")
from oracle import gen
from miplib_b200 import Engine

for cfg, graph in ((3, 0), (2, 0), (2, 1)):
    L = gen.config(cfg)
    with Engine(0) as e:
        e.load_lp(L)
        e.set_param("use_graph", graph)
        for pdl in (0, 1, 0, 1):
            e.set_param("use_pdl", pdl)
            e.solve()
            s = e.solve()
            row, _ = e.time_kernel(2, 30, 0)
            col, _ = e.time_kernel(3, 30, 0)
            print(f"C{cfg} graph {graph} pdl {pdl}: {s.status} {s.iterations} it, {s.iterations / s.iter_seconds:8.0f} it/s, "
                  f"in-solve row/col {s.row_kernel_us:.1f} / {s.col_kernel_us:.1f} us, back-to-back {row:.1f} / {col:.1f} us, "
                  f"obj {s.primal_obj:.9f}", flush=True)
