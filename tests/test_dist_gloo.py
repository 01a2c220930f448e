"""The N > 1 host path on CPU: two gloo ranks shard the replica axis, gather per-replica energies and
broadcast the best configuration (montecarlosolvers_b200.parallel) -- the only collective the path has."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, R, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from montecarlosolvers_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard(R, rank, world)
    rng = np.random.RandomState(123)
    all_e = rng.normal(size=R)
    all_c = (2 * rng.randint(2, size=(R, 11)) - 1).astype(np.int8)
    e, best, conf = parallel.gather_best(all_e[lo:hi], all_c[lo:hi], lo, R)
    ok = np.array_equal(e, all_e) and best == int(np.argmin(all_e)) and np.array_equal(conf, all_c[best])
    # device-resident form (torch tensors; CPU tensors under gloo, CUDA tensors under NCCL on the GPU box)
    import torch
    e2, best2, conf2 = parallel.gather_best_device(torch.from_numpy(all_e[lo:hi].copy()),
                                                   torch.from_numpy(all_c[lo:hi].copy()), lo, R)
    ok = ok and np.array_equal(e2.numpy(), all_e) and best2 == best and np.array_equal(conf2.numpy(), all_c[best])
    q.put((rank, bool(ok), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_best_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    R = 37  # uneven split: 19 + 18
    procs = [ctx.Process(target=_worker, args=(r, 2, port, R, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 19, 19, 37)


def test_single_process_gather_is_identity():
    from montecarlosolvers_b200 import parallel
    e = np.array([3.0, -1.0, 2.0])
    c = np.arange(6, dtype=np.int8).reshape(3, 2)
    ee, b, cc = parallel.gather_best(e, c, 0, 3)
    assert b == 1 and np.array_equal(cc, c[1]) and np.array_equal(ee, e)
    import torch
    e2, b2, c2 = parallel.gather_best_device(torch.from_numpy(e), torch.from_numpy(c), 0, 3)
    assert b2 == 1 and np.array_equal(c2.numpy(), c[1]) and np.array_equal(e2.numpy(), e)
