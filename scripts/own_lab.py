"""Two-GPU laboratory for the block-owner exchange (p2p.cu): what each piece of the protocol costs.

  gpurun --gpus 2 -- python scripts/own_lab.py [config] [world]

Times the row and the column kernel (30 back-to-back launches each, CUDA events, both ranks in
lock step) for the fused P2P exchange (1) and for the block-owner exchange (2) under every
MIPCU_OWN_VARIANT switch (with and without the NVSwitch multicast window): 1 no per-CTA publish fence, 2 relaxed polling + one fence,
4 no pushes (timing only), 8 no barrier (timing only).
"""
import os
import sys

import torch.multiprocessing as mp


def worker(rank, world, uid, cfg, out):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from miplib_b200 import Engine, sharding
    from oracle import gen
    L = gen.config(cfg)
    ptr, idx, val, lo, hi = sharding.shard_csr(L.A, L.row_lo, L.row_hi, rank, world)
    res = {}
    cases = [(1, 0, 0, 0, 1), (2, 1, 0, 1, 1), (2, 1, 0, 1, 2), (2, 1, 0, 1, 3), (2, 1, 0, 1, 4), (2, 13, 0, 0, 1), (2, 13, 0, 0, 2)]
    if os.environ.get("OWN_LAB_CASES"):
        cases = [tuple(int(v) for v in c.split(",")) for c in os.environ["OWN_LAB_CASES"].split(";")]
    with Engine(rank, rank, world, uid) as e:
        for ex, var, pdl, mc, blocks in cases:
            os.environ["MIPCU_OWN_VARIANT"] = str(var)
            os.environ["MIPCU_OWN_BLOCKS"] = str(blocks)
            e.set_param("shard_exchange", ex)
            e.set_param("use_pdl", pdl)
            e.set_param("multicast", mc)
            e.load(ptr, idx, val, L.c, L.lb, L.ub, lo, hi, L.maximize, n=L.n)
            mc = int(e.get_param("multicast")) if ex == 2 else 0
            e.time_kernel(2, 5)
            row, _ = e.time_kernel(2, 30)
            col, _ = e.time_kernel(3, 30)
            its = 0.0
            if var in (0, 1):           # complete protocol: a real solve
                e.solve()
                st = e.solve()
                its = st.iterations / st.iter_seconds
            res[(ex, var, pdl, mc, blocks)] = (row, col, its)
    out[rank] = res


if __name__ == "__main__":
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from miplib_b200 import Engine
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    uid = Engine.unique_id()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(worker, args=(world, uid, cfg, out), nprocs=world, join=True)
        res = dict(out)
    print("exchange variant pdl multicast(in use) blocks/CTA | row us (rank 0 / last) | col us (rank 0 / last) | solve it/s")
    for key in res[0]:
        r0, r1 = res[0][key], res[world - 1][key]
        print(f"{key[0]:8d} {key[1]:7d} {key[2]:3d} {key[3]:3d} {key[4]:3d} | {r0[0]:8.1f} {r1[0]:8.1f} | {r0[1]:8.1f} {r1[1]:8.1f} | {r0[2]:8.0f}")
