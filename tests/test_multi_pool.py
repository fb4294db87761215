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
    rows = pools.schedule_pools(_pools(), rank, world, dev.schedule, dist)
    np.save(os.path.join(out_dir, f"rows{rank}.npy"), rows)
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
    single = pools.schedule_pools(_pools(), 0, 1, emu_lib.emu_round().schedule)
    assert (r0 == single).all()
    assert (r0[:, 3] > 0).all()
