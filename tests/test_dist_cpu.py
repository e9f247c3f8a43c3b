"""world_size-2 gloo test (CPU) of the multi-GPU bookkeeping used by bench.py: row-block
sharding with no data-path collective, max-over-ranks timing, whole-job throughput."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aby3_b200 import distutil


def test_row_blocks_partition_the_rows():
    for total, world in [(4096, 1), (4096, 2), (4096, 8), (1000, 3), (7, 8)]:
        blocks = [distutil.row_block(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        for a, b in zip(blocks, blocks[1:]):
            assert a[1] == b[0]
        sizes = [b[1] - b[0] for b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, l = distutil.env_rank()
    assert (r, w, l) == (rank, world, rank)
    # each rank "processes" its own row block; the product of a sharded plaintext matmul must
    # equal the unsharded one without any exchange between ranks
    rng = np.random.default_rng(0)
    a = rng.integers(-2**62, 2**62, (64, 16), dtype=np.int64)
    b = rng.integers(-2**62, 2**62, (16, 8), dtype=np.int64)
    r0, r1 = distutil.row_block(64, rank, world)
    mine = a[r0:r1] @ b
    ms = 10.0 + 5.0 * rank                      # rank 1 is the slow one
    ms_max, launches, units = distutil.combine(dist, "cpu", ms, 7, float((r1 - r0) * 16 * 8))
    gathered = [torch.zeros(((64 // world), 8), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))       # test-only: check the shards tile the result
    if rank == 0:
        full = torch.cat(gathered).numpy()
        out.put((ms_max, launches, units, bool(np.array_equal(full, a @ b))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_combine_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ms_max, launches, units, ok = res
    assert ms_max == 15.0 and launches == 14 and units == 64 * 16 * 8 and ok
    assert distutil.throughput(units, ms_max) == pytest.approx(64 * 16 * 8 / 15e-3)
