"""N > 1 host logic on CPU with the gloo backend (world_size 2): batch sharding,
block-cyclic column ownership, and the info/retcode reduction.  No GPU involved."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import linearsolve_jl_b200 as ls
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. batch sharding: ranges partition [0, batch)
        for batch in (65536, 65537, 7, 1):
            a, b = ls.shard_batch(batch, rank, world)
            t = torch.tensor([a, b])
            allr = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
            dist.all_gather(allr, t)
            spans = sorted((int(x[0]), int(x[1])) for x in allr)
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
        # 2. block-cyclic columns: every global column owned exactly once, in block order
        for n, nb in ((65536, 256), (1000, 64), (300, 256), (17, 16)):
            cols = ls.block_cyclic_columns(n, nb, rank, world)
            cnt = torch.zeros(n, dtype=torch.long)
            cnt[torch.from_numpy(cols)] += 1
            dist.all_reduce(cnt)
            assert int(cnt.min()) == 1 and int(cnt.max()) == 1
            g = cols // nb
            assert np.all(g % world == rank) and np.all(np.diff(cols) > 0)
        # 3. retcode reduction: one rank sees a singular block -> the whole job fails
        info_local = torch.tensor([0 if rank == 0 else 5])
        infos = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(infos, info_local)
        assert ls.reduce_info([int(i) for i in infos]) == ls.ReturnCode.Failure
        assert ls.reduce_info([0, 0]) == ls.ReturnCode.Success
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=240) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
