"""Multi-process tests of the C++ drop-in header: N ranks = N processes, as ExaChem runs them (one MPI rank per GPU).
The TAMM stand-in's ProcGroup (oracle/shim/tamm/tamm.hpp) gives them rank / size, a barrier, a broadcast and the caller-side
sum over ranks (ccsd_t.cpp:262-263); the library's shared task counter (ccsdt_task_counter_open, the role of
AtomicCounterGA, ccsd_t_fused_driver.hpp:169-172,456) hands the tasks out; with CCSDT_B200_INTERNAL_ALLREDUCE the
library's own ncclAllReduce combines the energies (needs one GPU per rank)."""
import json
import os
import subprocess
import sys
import uuid

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_small.json")))
WORKER = os.path.join(ROOT, "tests", "multirank_worker.py")


def launch(nranks, mode, name, tmp_path, flavour="", env_extra=None, devices=None):
    if not os.path.exists(os.path.join(ROOT, "tests", "cpp", "_build", "libadapter_test_nccl.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    key = uuid.uuid4().hex[:12]
    seg = f"/dev/shm/tamm_shim_{key}"
    with open(seg, "wb") as f:                 # the launcher creates and zeroes the ranks' meeting point
        f.write(b"\0" * 4096)
    procs, outs = [], []
    try:
        for r in range(nranks):
            env = dict(os.environ, TAMM_SHIM_RANK=str(r), TAMM_SHIM_SIZE=str(nranks), TAMM_SHIM_KEY=key,
                       CCSDT_B200_COUNTER_KEY=key, CCSDT_B200_EXEC_TILESIZE="0", **(env_extra or {}))
            if devices is not None:
                env["WORKER_DEVICE"] = str(devices[r])
            out = str(tmp_path / f"rank{r}.json")
            outs.append(out)
            procs.append(subprocess.Popen([sys.executable, WORKER, mode, name, out, flavour], env=env))
        for p in procs:
            assert p.wait(timeout=300) == 0
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        os.unlink(seg)
    return [json.load(open(o)) for o in outs]


@pytest.mark.parametrize("nranks", [2, 3])
def test_procgroup_and_shared_task_counter_across_processes(nranks, tmp_path):
    n = 20000
    res = launch(nranks, "selftest", str(n), tmp_path)
    assert all(r["rc"] == 0 for r in res), res
    assert sum(r["out"][0] for r in res) == n                     # every ticket claimed exactly once ...
    assert sum(r["out"][1] for r in res) == n * (n - 1) // 2      # ... and they are 0..n-1
    for r in res:
        assert r["out"][2] == n and r["out"][3] == n * (n - 1) // 2 and r["out"][4] == 0x5eed1234


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["o3v9_ts4", "uhf_o3o2_v5v6_ts3"])
@pytest.mark.parametrize("dynamic", ["1", "0"])
def test_two_ranks_through_the_cpp_header_on_one_gpu(name, dynamic, tmp_path):
    """two processes call execute with rank 0 / 1 of 2 (sharing GPU 0): shared-counter hand-out (or the static split),
    rank partials returned, the caller's reduction gives the reference energy"""
    g = GOLD[name]
    res = launch(2, "execute", name, tmp_path, env_extra={"CCSDT_B200_DYNAMIC": dynamic})
    assert all(r["rc"] == 0 for r in res), res
    for r in res:
        assert abs(r["e1"] - float(g["energy1"])) <= 1e-9 and abs(r["e2"] - float(g["energy2"])) <= 1e-9
    assert sum(r["tasks_run"] for r in res) == len(g["tasks"])
    if dynamic == "0":
        assert all(r["tasks_run"] > 0 for r in res)


@pytest.mark.gpu
def test_node_shared_block_store_fetches_each_block_once_per_node(tmp_path):
    """two ranks with the shared directory (the adapter's default on one node): a block crosses PCIe once per NODE -- the
    rank that needs it second copies it from the first rank's HBM through CUDA IPC -- and the energies do not change"""
    name = "h2o_shape_ts7"
    g = GOLD[name]
    # static split in both runs, so that the two ranks need the same blocks with and without the shared store
    alone = launch(2, "execute", name, tmp_path, env_extra={"CCSDT_B200_SHARE": "0", "CCSDT_B200_DYNAMIC": "0"})
    shared = launch(2, "execute", name, tmp_path, env_extra={"CCSDT_B200_DYNAMIC": "0"})
    for r in alone + shared:
        assert r["rc"] == 0, r
        assert abs(r["e1"] - float(g["energy1"])) <= 1e-9 and abs(r["e2"] - float(g["energy2"])) <= 1e-9
    # (not bit-identical: with the shared counter the ranks' partial sums group the tasks differently from run to run)
    assert all(r["blocks_from_peers"] == 0 for r in alone)
    host_alone, host_shared = sum(r["blocks_fetched"] for r in alone), sum(r["blocks_fetched"] for r in shared)
    peers = sum(r["blocks_from_peers"] for r in shared)
    assert peers > 0 and host_shared < host_alone
    assert host_shared + peers == host_alone                          # every duplicate fetch became a peer copy ...
    assert all(r["gets"] == r["blocks_fetched"] for r in shared)      # ... and Tensor::get saw only the host fetches


@pytest.mark.gpu
def test_internal_nccl_allreduce_through_the_cpp_header(tmp_path):
    """CCSDT_B200_INTERNAL_ALLREDUCE: the library's ncclAllReduce inside execute (one GPU per rank); rank 0 returns the
    total and the others 0, so the caller's reduction still yields the total"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    g = GOLD["o3v9_ts4"]
    res = launch(2, "execute", "o3v9_ts4", tmp_path, flavour="_nccl", devices=[0, 1])
    assert all(r["rc"] == 0 for r in res), res
    for r in res:
        assert abs(r["e1"] - float(g["energy1"])) <= 1e-9 and abs(r["e2"] - float(g["energy2"])) <= 1e-9


@pytest.mark.gpu
def test_node_shared_block_store_evicts_under_a_small_budget(tmp_path):
    """the shared store with 64 KB of HBM per rank (the working set is larger): blocks are evicted -- never while a peer is
    copying them --, their directory entries are tombstoned and reused, evicted blocks come back from the host or from the
    other rank, and the energies stay within the bar"""
    name = "h2o_shape_ts7"
    g = GOLD[name]
    res = launch(2, "execute", name, tmp_path, env_extra={"CCSDT_B200_DYNAMIC": "0", "CCSDT_B200_BLOCK_BUDGET_MB": "0.0625"})
    for r in res:
        assert r["rc"] == 0, r
        assert abs(r["e1"] - float(g["energy1"])) <= 1e-9 and abs(r["e2"] - float(g["energy2"])) <= 1e-9
    assert sum(r["blocks_evicted"] for r in res) > 0

