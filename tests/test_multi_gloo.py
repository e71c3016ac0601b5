"""Multi-rank host logic on CPU (gloo, world_size 2): window sharding, the table
broadcast and the gather of results.  The per-rank "demod" here is the CPU oracle
standing in for the GPU kernel -- this test is about the N>1 plumbing only.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_window_range_partitions_everything():
    import lora_sdr_b200  # noqa: F401
    from lora_sdr_b200 import sharding

    for n in (0, 1, 7, 64, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.window_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [e[1] - e[0] for e in edges]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.channel_range(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        sharding.window_range(10, 2, 2)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import lora_sdr_b200 as m
    from lora_sdr_b200 import sharding
    from oracle import pyoracle
    from util import synth_windows

    dist.init_process_group("gloo", rank=rank, world_size=world)
    sf, nw = 7, 101
    # rank 0 owns the tables; the others start from garbage and receive the broadcast
    tabs = []
    for which in (m.TABLE_UP, m.TABLE_TWIDDLE):
        t = torch.from_numpy(m.build_table(sf, which).view(np.float32).copy())
        if rank != 0:
            t.zero_()
        tabs.append(t)
    sharding.broadcast_tables(tabs, src=0)
    up = tabs[0].numpy().view(np.complex64)
    assert up.tobytes() == m.build_table(sf, m.TABLE_UP).tobytes()

    iq, _ = synth_windows(sf, nw, seed=5, snr_db=0.0)       # same bytes on every rank
    w0, w1 = sharding.window_range(nw, rank, world)
    orc = pyoracle.Oracle()
    res, _, _, _ = orc.demod_windows(sf, iq[w0:w1], w1 - w0, up)
    # pad to a common length for all_gather, then trim
    pad = (nw + world - 1) // world
    loc = torch.full((pad,), -1, dtype=torch.int32)
    loc[: w1 - w0] = torch.from_numpy(res["index"].copy())
    parts = sharding.gather_results(loc)
    merged = np.concatenate([p.numpy()[: sharding.window_range(nw, r, world)[1] - sharding.window_range(nw, r, world)[0]]
                             for r, p in enumerate(parts)])
    full, _, _, _ = orc.demod_windows(sf, iq, nw, up)
    assert np.array_equal(merged, full["index"])
    if rank == 0:
        Path(tmp, "ok").write_text("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    import torch.multiprocessing as mp

    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"
