"""Host-side logic of tensor parallelism on CPU: gloo, world_size 2 (no GPU, no compute calls)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctranslate2_b200 import parallel  # noqa: E402
from oracle import ct2_oracle as O  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, s = parallel.default_rank_and_size(None, None)
        assert (r, s) == (rank, world)
        mine = bytes([rank]) * 64                       # stands for the rank's cudaIpcMemHandle
        handles = parallel.exchange_handles(mine, r, s)
        assert handles == [bytes([k]) * 64 for k in range(world)]
        # the partition every rank derives covers each dimension exactly once
        H, Hkv, D, F = 8, 4, 16, 256
        rows = O.tp_qkv_rows(H, Hkv, D, r, s)
        gathered = [None] * world
        dist.all_gather_object(gathered, rows.tolist())
        assert sorted(sum(gathered, [])) == list(range((H + 2 * Hkv) * D))
        assert parallel.shard_range(F, r, s) == O.tp_shard_rows(F, r, s)
        # a mismatching rank/size is rejected before anything is exchanged
        try:
            parallel.exchange_handles(mine, r, s + 1)
            raise AssertionError("expected ValueError")
        except ValueError:
            pass
        q.put((rank, "ok"))
    except Exception as e:                              # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_handle_exchange_and_partition_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
    assert results == {0: "ok", 1: "ok"}, results


def test_shard_range_rejects_indivisible():
    with pytest.raises(ValueError):
        parallel.shard_range(10, 0, 4)
    assert parallel.shard_range(8, 1, 4) == (2, 4)
    with pytest.raises(RuntimeError):
        parallel.default_rank_and_size(None, None)      # torch.distributed is not initialised here
