"""Multi-rank host logic on CPU: LPT sharding, max-over-ranks timing and the final gather of
fitted hyper-parameters, with the gloo backend and world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from medgp_b200 import shard


def test_lpt_balances_cubic_load():
    rng = np.random.default_rng(0)
    sizes = rng.integers(300, 1500, 512)
    for world in (2, 4, 8):
        a = shard.lpt_assign(sizes, world)
        load = np.array([(sizes[a == r].astype(float) ** 3).sum() for r in range(world)])
        assert set(a) == set(range(world))
        assert load.max() / load.min() < 1.02          # within 2 % of perfect balance
    assert (shard.lpt_assign([10, 20, 30], 1) == 0).all()


def test_cpp_lpt_equals_python_lpt(tmp_path):
    """the front-ends' C++ deal (medgp_lpt_assign) is the Python one (shard.lpt_assign), patient for
    patient, and balances the cubic load of a C3-like cohort"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle"), "host_on_oracle"], check=True, capture_output=True)
    exe = os.path.join(root, "oracle", "_build", "host_check")
    sizes = np.random.default_rng(3).integers(300, 1501, 600)
    for world in (1, 2, 8):
        out = subprocess.run([exe, "lpt", str(world)] + [str(int(v)) for v in sizes], check=True, capture_output=True, text=True).stdout
        got = np.array([int(line.split()[1]) for line in out.split("\n") if line.startswith("s ")])
        want = shard.lpt_assign(sizes, world)
        assert np.array_equal(got, want)
        load = np.array([(sizes[got == r].astype(float) ** 3).sum() for r in range(world)])
        assert load.max() / load.mean() < 1.01


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = np.arange(10, 30)
    owner = np.nonzero(shard.lpt_assign(sizes, world) == rank)[0]
    theta_local = np.stack([np.full(5, float(i)) for i in owner]) if len(owner) else np.zeros((0, 5))
    ms, evals = shard.reduce_report(dist, torch.device("cpu"), 10.0 + 5.0 * rank, 100 * (rank + 1))
    full = shard.gather_theta(dist, torch.device("cpu"), theta_local, owner, len(sizes))
    q.put((rank, ms, evals, full))
    dist.destroy_process_group()


def test_two_rank_reduce_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, evals, full in results:
        assert ms == 15.0                      # max over ranks, never the local time
        assert evals == 300                    # whole-job total
        assert np.array_equal(full[:, 0], np.arange(20, dtype=float))   # every row from its owner
