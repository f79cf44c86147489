"""Multi-GPU plumbing: contiguous output row bands per rank and the gather of the
finished bands.  One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards into fully independent units: every output pixel depends only on (seed, offsets,
lambda in a bounded neighbourhood), and grains are pure functions of (seed, i, j, lambda(cell))
(src/rng.rs:26-34, src/pixelwise.rs:68-82).  Each rank therefore renders its band with the margin
cells regenerated locally -- there is no halo exchange and no data-path collective; the only
communication is the gather of the final image.
"""
from __future__ import annotations


def band_rows(out_h: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [begin, end) of `rank`: contiguous, disjoint, covering [0, out_h), sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(out_h, world)
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    return begin, end


def max_band_rows(out_h: int, world: int) -> int:
    return (out_h + world - 1) // world


_gather_cache: dict = {}


def gather_bands(band, out_h: int, rank: int, world: int, dst: int = 0, group=None):
    """Gather per-rank bands [..., rows_r, W(, C)] (row axis = -2 for planes [P,rows,W]; pass tensors whose
    dim 0 is the row axis) to `dst`.  Bands are padded to the common maximum so one collective moves
    everything; returns the assembled [out_h, ...] tensor on dst, None elsewhere.  The receive buffer is
    allocated once per shape and reused (the returned tensor is overwritten by the next call); when the
    rows divide evenly the bands land directly in place, otherwise they are compacted with one copy."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return band
    mx = max_band_rows(out_h, world)
    rows = band.shape[0]
    if rows < mx:
        pad = torch.zeros((mx - rows,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
        band = torch.cat([band, pad], dim=0)
    band = band.contiguous()
    if rank == dst:
        key = (tuple(band.shape), band.dtype, band.device, world)
        recv = _gather_cache.get(key)
        if recv is None:
            recv = torch.empty((world * mx,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
            _gather_cache.clear()
            _gather_cache[key] = recv
        parts = [recv[r * mx:(r + 1) * mx] for r in range(world)]
        dist.gather(band, parts, dst=dst, group=group)
        if out_h == world * mx:
            return recv
        pieces = []
        for r in range(world):
            b, e = band_rows(out_h, r, world)
            pieces.append(parts[r][: e - b])
        return torch.cat(pieces, dim=0)
    dist.gather(band, None, dst=dst, group=group)
    return None


class PeerImage:
    """The finished image on `dst`, mapped into every rank's address space over NVLink / NVSwitch.

    Bands are independent, so the only inter-GPU step of the path is "the finished bands end up on
    GPU `dst`".  Instead of rendering into a local buffer and gathering afterwards, every rank hands
    the engine a pointer INTO `dst`'s image: the kernels' own coalesced stores of the band rows travel
    over NVLink as peer writes while the band is being computed, and the step ends with one
    device-side barrier on the engine's stream.  There is no staging copy and no collective.

    The mapping is torch symmetric memory (cuMem allocations whose handles torch exchanges between the
    ranks; the barrier runs through its signal pads).  `PeerImage.create` returns None when it cannot
    be set up on every rank (the caller then uses `gather_bands`).
    """

    def __init__(self, mode, local, target, finish, keep):
        self.mode, self.local, self.target, self._finish, self._keep = mode, local, target, finish, keep

    def finish(self):
        """Barrier on the current CUDA stream: after it, every rank's band is complete in `dst`'s image."""
        self._finish()

    @staticmethod
    def _agree(ok: bool, device, group) -> bool:
        import torch
        import torch.distributed as dist
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return bool(int(flag.item()))

    @classmethod
    def create(cls, shape, dtype, device, rank: int, world: int, dst: int = 0, group=None, modes=("symm",)):
        import torch
        import torch.distributed as dist

        if world == 1:
            return None
        shape = tuple(int(s) for s in shape)
        for mode in modes:
            made, err = None, None
            try:
                if mode == "symm":
                    import torch.distributed._symmetric_memory as symm
                    local = symm.empty(shape, dtype=dtype, device=device)
                    hdl = symm.rendezvous(local, group if group is not None else dist.group.WORLD)
                    target = local if rank == dst else hdl.get_buffer(dst, shape, dtype)
                    made = cls(mode, local, target, hdl.barrier, (hdl,))
                else:
                    raise ValueError(f"unknown peer image mode {mode!r}")
            except Exception as e:  # set-up only: a mode that cannot be established is skipped on ALL ranks
                err = e
            if cls._agree(made is not None, device, group):
                return made
            del made
            if err is not None and rank == dst:
                import sys
                print(f"[film_grain_b200.dist] peer image mode {mode!r} unavailable: {err!r}", file=sys.stderr)
        return None


class SharedHostImage:
    """The finished image in page-locked HOST memory shared by every rank of the box.

    One /dev/shm mapping, registered with cudaHostRegister (portable, mapped) in every process.  Each rank
    passes the planes of this block as the `out` buffers of the host-pointer C-ABI call
    (`fg_render_planes` with its row band): the engine recognises page-locked output and lets the kernels
    store the band rows straight into it, so every GPU delivers its band over its own PCIe link while it
    renders and nothing funnels through GPU 0.  After a barrier the whole image is in host memory.
    `create` returns None when the mapping cannot be set up on every rank.
    """

    def __init__(self, array, path, owner):
        self.array, self._path, self._owner = array, path, owner

    @classmethod
    def create(cls, shape, rank: int, world: int, device, group=None):
        import os
        import uuid

        import numpy as np
        import torch
        import torch.distributed as dist

        tag = [uuid.uuid4().hex if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(tag, src=0, group=group)
        path = f"/dev/shm/fg_b200_{tag[0]}"
        made, err = None, None
        try:
            if rank == 0:
                arr = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=tuple(shape))
                arr[...] = 0.0
            if world > 1:
                dist.barrier(group=group)
            if rank != 0:
                arr = np.load(path, mmap_mode="r+")
            rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 1 | 2)  # portable | mapped
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister -> {rc}")
            made = cls(arr, path, rank == 0)
        except Exception as e:  # set-up only
            err = e
        ok = PeerImage._agree(made is not None, device, group) if world > 1 else made is not None
        if not ok:
            if made is not None:
                made.close()
            if err is not None and rank == 0:
                import sys
                print(f"[film_grain_b200.dist] shared host image unavailable: {err!r}", file=sys.stderr)
            if rank == 0 and os.path.exists(path):
                os.unlink(path)
            return None
        return made

    def close(self):
        import os

        import torch
        try:
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
        except Exception:
            pass
        if self._owner and os.path.exists(self._path):
            os.unlink(self._path)
