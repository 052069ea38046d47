"""A/B of MIPCU_PARAM_ROW_SORT = 1 / 2 / 3 / 4 on one box: isolated launch times of the two fused kernels (config 3)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from miplib_b200 import Engine
from oracle import gen
L = gen.config(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
for rep in range(2):
    for mode in (1, 2, 3, 4):
        with Engine(0) as e:
            e.set_param("row_sort", mode)
            e.load_lp(L)
            e.prepare()
            e.time_kernel(4, 100, False)
            row, _ = e.time_kernel(2, 300, False)
            col, _ = e.time_kernel(3, 300, False)
            it, _ = e.time_kernel(4, 1000, False)
        print(json.dumps({"row_sort": mode, "row_step_us": row, "col_step_us": col, "iteration_us_1000": it}), flush=True)
