"""N>1 host logic on CPU: the round-robin tile partition and the all-gather assembly of the frame
(world_size 2 and 3, gloo backend).  The GPU path runs the same TileExchange code over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpuvoxelraycaster_b200.frame import TileExchange
    ex = TileExchange(W, H, rank, world, torch.device("cpu"))
    # the "rendered" frame: a known image; each rank only fills the rows it owns, the rest stays poisoned
    rng = np.random.default_rng(123)
    full = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    mine = np.full((ex.H_pad, W, 4), 0xAB, np.uint8)
    rows = ex.owned_rows()
    mine[rows] = full[rows]
    out = ex.gather(torch.from_numpy(mine).reshape(-1))
    got = out.view(ex.H_pad, W, 4)[:H].numpy()
    ok = np.array_equal(got, full)
    # every row is owned exactly once across ranks
    counts = torch.zeros(H, dtype=torch.int32)
    counts[rows] = 1
    dist.all_reduce(counts)
    ok = ok and bool((counts == 1).all())
    open(os.path.join(result_dir, "rank%d" % rank), "w").write("ok" if ok else "fail")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H", [(2, 64, 36), (2, 33, 10), (3, 40, 27)])
def test_tile_partition_and_gather_gloo(tmp_path, world, W, H):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, W, H, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("rank%d" % r)).read() == "ok"


def test_single_rank_is_identity():
    from cpuvoxelraycaster_b200.frame import TileExchange
    ex = TileExchange(16, 9, 0, 1, torch.device("cpu"))
    assert ex.owned_rows() == list(range(9)) and ex.H_pad == 12
    t = torch.arange(ex.H_pad * 16 * 4, dtype=torch.int64).to(torch.uint8)
    assert ex.gather(t) is t
