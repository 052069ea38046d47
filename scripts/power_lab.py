"""Board power and SM clock per kernel variant: config 3, the shipped iteration
(row step + col step back to back) run for a few seconds per variant while nvidia-smi samples
power.draw / clocks.sm every 100 ms.  Which variant is fastest IN A LONG RUN (under the 1 kW cap), not
in a 30-launch burst.   usage: python scripts/power_lab.py > gpurun_out/power_lab.json"""
import json
import subprocess
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from miplib_b200 import Engine
from oracle import gen


class Sampler:
    def __init__(self):
        self.rows = []
        self.proc = subprocess.Popen(
            ["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks.mem,temperature.gpu,"
             "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"],
            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for ln in self.proc.stdout:
            f = [t.strip() for t in ln.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(f[0]), float(f[1]), float(f[2]), float(f[3]), f[4]))
            except (ValueError, IndexError):
                pass

    def window(self, t0, t1):
        w = [r for r in self.rows if t0 + 0.5 <= r[0] <= t1]
        if not w:
            return {}
        return {"sm_mhz_median": float(np.median([r[1] for r in w])), "power_w_median": float(np.median([r[2] for r in w])),
                "power_w_max": float(max(r[2] for r in w)), "mem_mhz": float(np.median([r[3] for r in w])),
                "temp_c_max": float(max(r[4] for r in w)), "power_cap_active": sum(r[5].lower().startswith("active") for r in w),
                "samples": len(w)}


def main():
    L = gen.config(3)
    variants = [("SELL-32-256 lane per row (shipped)", {}),
                ("CSR, 8 lanes per row (round-1 v1)", {"threads_per_row": 8}),
                ("CSR, 4 lanes per row", {"threads_per_row": 4}),
                ("TMA-streamed tiles", {"use_tma": 1}),
                ("plain SpMV pair (K v, K' v; no step fused)", {"plain": 1})]
    smp = Sampler()
    out = []
    for name, params in variants:
        with Engine(0) as e:
            plain = params.pop("plain", 0)
            for k, v in params.items():
                e.set_param(k, v)
            e.load_lp(L)
            e.prepare()
            e.time_kernel(4, 50, False)
            time.sleep(1.0)          # let the board fall back to idle between variants
            t0 = time.perf_counter()
            us_all = []
            while time.perf_counter() - t0 < 4.0:
                if plain:
                    a, _ = e.time_kernel(0, 500, False)
                    b, _ = e.time_kernel(1, 500, False)
                    us_all.append(a + b)
                else:
                    us, _ = e.time_kernel(4, 1000, False)
                    us_all.append(us)
            t1 = time.perf_counter()
            burst, _ = (e.time_kernel(4, 30, False) if not plain else (None, None))
        rec = {"variant": name, "iteration_us_sustained": float(np.mean(us_all[1:] or us_all)),
               "iteration_us_first_block": float(us_all[0]), **smp.window(t0, t1)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    smp.proc.terminate()


if __name__ == "__main__":
    main()
