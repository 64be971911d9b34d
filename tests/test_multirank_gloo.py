"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups.  Each rank takes the tasks the
library's static split (ccsdt_partition through the C ABI) gives it, evaluates them with the oracle as
the stand-in for the GPU kernel (this is a test: the oracle is the checker), and the rank partials are
combined by the path's one collective (exachem_b200.multigpu.combine_energies)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from exachem_b200 import driver as drv, multigpu, synthetic as syn

CFG = dict(oa=4, ob=4, va=7, vb=7, ts=3, seed=5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        orc = Oracle()
        c = CFG
        sp, osp = drv.setup_mo_space(c["oa"], c["ob"], c["va"], c["vb"], c["ts"]), orc.tiles(c["oa"], c["ob"], c["va"], c["vb"], c["ts"])
        T = syn.dense_all(syn.Orbitals(c["oa"], c["ob"], c["va"], c["vb"]), c["seed"])
        info = multigpu.RankInfo.from_env()
        assert (info.rank, info.nranks) == (rank, world)
        own = drv.partition(sp, True, world)
        _, _, per_task = orc.run(osp, T, True, per_task=True)
        mine = own == rank
        e1, e2 = float(per_task[mine, 0].sum()), float(per_task[mine, 1].sum())
        t1, t2 = multigpu.combine_energies(e1, e2)
        p1, p2 = multigpu.caller_side_partials(t1, t2, rank)
        q.put((rank, int(mine.sum()), e1, e2, t1, t2, p1, p2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rank_partials_all_reduce_to_the_oracle_total(orc, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    c = CFG
    osp = orc.tiles(c["oa"], c["ob"], c["va"], c["vb"], c["ts"])
    T = syn.dense_all(syn.Orbitals(c["oa"], c["ob"], c["va"], c["vb"]), c["seed"])
    ref1, ref2 = orc.run(osp, T, True)[:2]
    n_tasks = len(orc.enumerate(osp, True)[0])
    assert sum(r[1] for r in res) == n_tasks and all(r[1] > 0 for r in res)      # complete, nobody idle
    for r in res:
        assert abs(r[4] - ref1) < 1e-12 and abs(r[5] - ref2) < 1e-12           # every rank holds the total
    # what the ExaChem caller would reduce (ccsd_t.cpp:262-263) still sums to the total exactly once
    assert abs(sum(r[6] for r in res) - ref1) < 1e-12 and abs(sum(r[7] for r in res) - ref2) < 1e-12
