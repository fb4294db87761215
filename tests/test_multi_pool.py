"""N>1 path on CPU: two processes (torch.distributed, gloo) split the pools of a cycle between
them exactly like bench.py / the GPU box does with NCCL, and every rank ends up with the same
per-pool summary a single process computes.  The per-pool rounds run through the emulated kernel
build (tests/emu_lib.py) so no GPU is needed."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pools():
    from armada_b200 import synth
    return [synth.random_round(200 + i, n_nodes=60, n_queues=4, n_jobs=260, n_running=60).to_input() for i in range(5)]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import emu_lib
    from armada_b200 import pools
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = emu_lib.emu_round()
    rows, claims = pools.schedule_pools(_pools(), rank, world, dev.schedule, dist)
    np.save(os.path.join(out_dir, f"rows{rank}.npy"), rows)
    np.save(os.path.join(out_dir, f"claims{rank}.npy"), claims)
    dist.destroy_process_group()


def test_pools_are_split_over_two_ranks(tmp_path):
    sys.path.insert(0, ROOT)
    from armada_b200 import pools
    import emu_lib
    assert pools.pools_of_rank(5, 0, 2) == [0, 2, 4] and pools.pools_of_rank(5, 1, 2) == [1, 3]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rows0.npy")
    r1 = np.load(tmp_path / "rows1.npy")
    assert (r0 == r1).all()
    c0 = np.load(tmp_path / "claims0.npy")
    c1 = np.load(tmp_path / "claims1.npy")
    assert (c0 == c1).all()  # one all_gather leaves the whole cycle's claims on every rank
    single, sclaims = pools.schedule_pools(_pools(), 0, 1, emu_lib.emu_round().schedule)
    assert (r0 == single).all()
    assert (c0 == sclaims).all()
    assert (r0[:, 3] > 0).all()
    assert ((c0 != 0xFFFFFFFF).sum(axis=1) == r0[:, 1]).all()  # claims per pool == jobs scheduled


def test_pipelined_cycle_equals_pool_by_pool():
    """PoolCycle.schedule_cycle (upload of pool k+1 on a worker thread under the run of pool k, two
    device contexts) returns exactly what scheduling the pools one by one returns."""
    sys.path.insert(0, ROOT)
    from armada_b200 import pools
    import emu_lib
    inputs = _pools()
    cyc = pools.PoolCycle(inputs, 0, 1, emu_lib.emu_round)
    got = cyc.schedule_cycle()
    cyc.upload_resident()
    stats = cyc.run_resident()
    one = emu_lib.emu_round()
    for i, p in enumerate(cyc.mine):
        want = one.schedule(inputs[p])
        assert not got[p].diff(want)
        assert not cyc.download_resident(i).diff(want)
        assert int(stats[i].placements) == int(want.stats.placements)
    cyc.close()
