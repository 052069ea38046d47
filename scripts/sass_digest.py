"""SASS digest of the shipping kernels in miplib_b200/libmiplib_cuda.so (no GPU needed):
registers, spills, shared memory (cuobjdump -res-usage) and the load / store mnemonics each kernel
contains (cuobjdump -sass), written as a Markdown table.  `python scripts/sass_digest.py > profiles/r02_sass_digest.md`"""
import collections, re, subprocess, sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "miplib_b200" / "libmiplib_cuda.so"
WANT = ["k_row_step<0, false>", "k_row_step<0, true>", "k_col_step<0, false, false>", "k_col_step<0, true, true>",
        "k_kkt_row<0>", "k_spmv<0>", "k_row_step<8, false>", "k_col_step<8, false, false>",
        "k_row_step_own<0, false, false>", "k_row_step_own<0, false, true>", "k_col_step_own<0, false, false, false>",
        "k_col_step_own<0, false, false, true>", "k_primal_step_p2p<false, false>",
        "k_row_step_batch_fed<4>", "k_row_step_batch_fed<1>", "k_col_step_batch<false, false, 2>",
        "k_col_step_batch<false, false, 1>", "k_row_step_batch<true, 2>", "k_col_step_batch<true, true, 2>",
        "k_row_batch_long<false, 0, 2>", "k_publish_batch",
        "k_ctrl_check", "k_ctrl_check_batch", "k_restart_apply"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in ln:
        usage[cur] = dict(re.findall(r"(\w+):(\d+)", ln))
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
ops = collections.defaultdict(collections.Counter)
cur = None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if cur and m:
        op = m.group(1)
        if op.startswith(("LDG", "STG", "LDS", "STS", "LDL", "STL", "RED", "ATOM", "SHFL", "BAR", "DFMA", "UBLKCP", "UTMALDG", "SYNCS", "MEMBAR", "CCTL", "ERRBAR")):
            ops[cur][op] += 1
names = demangle(list(usage))
print("# Round 2 - SASS digest of the shipping kernels (`scripts/sass_digest.py`, cuobjdump on the sm_100a cubin)\n")
print("Spills = `LDL` / `STL` instructions and the `STACK` bytes ptxas reserved.  Matrix streams are `LDG.E.(64.)CONSTANT` with an")
print("`.EF`/no-allocate hint only where the source asks; gathers are `LDG.E.64.CONSTANT`; no kernel uses tensor, TMA or")
print("texture instructions on the default path (the opt-in TMA variant shows `UBLKCP` + `SYNCS`).\n")
print("| kernel | registers | stack (spill) bytes | static smem | loads / stores / others |")
print("|---|---|---|---|---|")
for want in WANT:
    hits = [k for k, v in names.items() if want in v and "mipcu::" in v]
    for k in hits[:1]:
        u = usage[k]
        o = ops.get(k, {})
        line = ", ".join(f"{op} x{c}" for op, c in sorted(o.items()))
        print(f"| `{want}` | {u.get('REG')} | {u.get('STACK')} | {u.get('SHARED')} | {line} |")
tot_spill = sum(1 for k in usage if int(usage[k].get("STACK", 0)) > 0 and "mipcu" in names.get(k, ""))
print(f"\nKernels of the library: {sum(1 for k in usage if 'mipcu' in names.get(k, ''))}; with a non-zero stack frame: {tot_spill}.")
