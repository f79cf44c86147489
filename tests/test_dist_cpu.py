"""Host-side multi-GPU logic on CPU: band partition properties and a world_size-2 gloo gather."""
import os
import socket

import numpy as np
import pytest

from film_grain_b200.dist import band_rows, gather_bands, max_band_rows


@pytest.mark.parametrize("h", [1, 2, 7, 270, 2160, 2161, 8192])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_band_rows_partition(h, world):
    bands = [band_rows(h, r, world) for r in range(world)]
    assert bands[0][0] == 0 and bands[-1][1] == h
    for (a, b), (c, d) in zip(bands[:-1], bands[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in bands]
    assert max(sizes) - min(sizes) <= 1 and max(sizes) == max_band_rows(h, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, h, w, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(h * w * 3, dtype=torch.float32).reshape(h, w, 3)
        b, e = band_rows(h, rank, world)
        out = gather_bands(full[b:e].clone(), h, rank, world)
        if rank == 0:
            q.put(bool(torch.equal(out, full)))
        else:
            q.put(out is None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("h", [9, 10])
def test_gather_bands_gloo_world2(h):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, h, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(res)


def test_peer_image_is_not_used_for_a_single_rank():
    """world == 1: no peer mapping, the engine renders into its own buffer (and importing the module needs no GPU)."""
    import torch
    from film_grain_b200.dist import PeerImage
    assert PeerImage.create((3, 8, 8), torch.float32, "cpu", 0, 1) is None
