"""CPU tier, N > 1 path: two gloo ranks shard a receiver bank, each runs ITS receivers, rank 0 gathers the
spectrum rows and the max-over-ranks timing.  The compute engine here is the host emulation of the
kernel phases (tests/devtools, CPU-only logic check) because the CUDA library needs a GPU; the sharding,
gather and reduction code is the code bench.py runs under NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import oracle_py as O
import rx_driver
from t41_sdr_b200 import rx, sharding


def test_shard_ranges_cover_and_are_disjoint():
    for n in (0, 1, 7, 1024, 8191):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = sharding.shard_range(n, r, world)
                seen.extend(range(first, first + count))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


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
        case = cases.c4_zoom_rows(n=5, T=3)           # 5 receivers over 2 ranks: ragged shards (2 + 3)
        first, count = sharding.shard_range(case.n_streams, rank, world)
        eng = rx_driver.EmulReceiver(count)
        plist = [rx_driver.to_rx_params(p) for p in case.segments[0][0][first:first + count]]
        eng.set_params_each(plist)
        iq = np.stack(case.iq[first:first + count])
        out = eng.process(iq, row_every=case.row_every, flags=rx.FLAG_EXACT_NCO)
        eng.close()
        rows = sharding.gather_rows(torch.from_numpy(out["spec"]), case.n_streams, dst=0)
        slowest = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            q.put((rows.numpy(), slowest))
        else:
            assert rows is None
    except Exception as e:  # surface worker failures instead of letting the parent time out
        if rank == 0:
            q.put(("error", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_gather_and_reduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, slowest = q.get(timeout=120)
    assert not isinstance(rows, str), slowest
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert slowest == 11.0
    case = cases.c4_zoom_rows(n=5, T=3)
    want = cases.run_case_on(case, lambda p: O.OracleStream(p))
    assert rows.shape == (5, 3, 512)
    for s in range(5):
        assert np.array_equal(rows[s], want[s]["spec"]), s
