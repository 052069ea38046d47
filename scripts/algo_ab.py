"""GPU side of the step-rule A/B: what one iteration costs as shipped
(constant step, reflected Halpern, two launches) and with the per-iteration machinery an adaptive
step-size rule needs (three reductions in both epilogues + a single-CTA controller launch between
iterations), on configs 2 and 3; and the whole solve for reference.  Iteration counts of the two
methods are CPU facts (oracle prototype; HiGHS' PDLP in tests/golden/objectives.json).
    python scripts/algo_ab.py > gpurun_out/algo_ab.json"""
import json, sys
sys.path.insert(0, ".")
from oracle import gen
from miplib_b200 import Engine

out = {}
for cfg in (2, 3):
    L = gen.config(cfg)
    with Engine(0) as e:
        e.load_lp(L)
        e.set_param("use_graph", 0)
        e.prepare()
        plain, _ = e.time_kernel(4, 200, False)
        adapt, _ = e.time_kernel(5, 200, False)
        e.set_param("use_graph", -1)
        st = e.solve()
        st = e.solve()
        out[f"C{cfg}"] = {"iteration_us_constant_step": plain, "iteration_us_with_adaptive_rule_overhead": adapt,
                          "overhead_ratio": adapt / plain, "solve_iterations": st.iterations,
                          "solve_seconds": st.solve_seconds, "iter_seconds": st.iter_seconds}
    print(f"C{cfg}", out[f"C{cfg}"], file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
