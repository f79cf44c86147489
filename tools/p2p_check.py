#!/usr/bin/env python
"""Multi-GPU band check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py

Every rank renders its row band (a) straight into GPU 0's peer-mapped image (PeerImage, each mode that
can be set up), (b) locally + NCCL gather and (c) through the host-pointer call into one page-locked host image
shared by all ranks; rank 0 renders the whole image alone and requires both
assembled images to be bitwise identical to it.  Prints one line per check; exit code 1 on a mismatch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import film_grain_b200 as fg
    from film_grain_b200 import host as H
    from film_grain_b200.dist import PeerImage, SharedHostImage, band_rows, gather_bands

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    ctx = fg.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for algo_name, w, h, radius, n, zoom in (("pixel", 1000, 701, 0.1, 48, 1.0), ("pixel", 300, 200, 0.05, 32, 2.5),
                                              ("grain", 640, 403, 0.5, 40, 1.0)):
        img = np.random.default_rng(7).integers(0, 256, (h, w, 3), dtype=np.uint8)
        params = H.ParamsBuilder(radius_mean=radius, n_samples=n, zoom=zoom,
                                 algo=H.Algo.Pixel if algo_name == "pixel" else H.Algo.Grain, color_mode=H.ColorMode.Rgb).build()
        d = H.derive_common(params, (w, h))
        lam = np.stack([H.lambda_plane((img[:, :, c].astype(np.float32) / np.float32(255.0)).astype(np.float32), d.inv_e_pi_r2)
                        for c in range(3)])
        offsets = d.offsets_input if algo_name == "pixel" else d.offsets
        algo = fg.FG_ALGO_PIXEL if algo_name == "pixel" else fg.FG_ALGO_GRAIN
        out_w, out_h = d.output_width, d.output_height
        rb, re = band_rows(out_h, rank, world)
        blk = H._band(d.block, (rb, re))
        if algo_name == "grain":
            blk.path = fg.FG_PATH_STAGED  # the shared-memory tile rasteriser (AUTO would pick the global mask for so few tiles)
        d_lam = torch.from_numpy(lam).to(dev)
        d_off = torch.from_numpy(np.ascontiguousarray(offsets)).to(dev)
        ref = None
        if rank == 0:
            ref = torch.zeros((3, out_h, out_w), dtype=torch.float32, device=dev)
            ctx.render_planes_device(H._band(d.block, None), algo, 3, d_lam.data_ptr(), d_off.data_ptr(), ref.data_ptr(), sync=True)
        with torch.cuda.stream(stream):
            for mode in ("symm",):
                peer = PeerImage.create((3, out_h, out_w), torch.float32, dev, rank, world, modes=(mode,))
                if peer is None:
                    if rank == 0:
                        print(f"{algo_name} {w}x{h} zoom {zoom}: peer mode {mode}: NOT AVAILABLE", flush=True)
                    continue
                peer.local.fill_(-1.0)
                peer.finish()
                for _ in range(2):
                    ctx.render_planes_device(blk, algo, 3, d_lam.data_ptr(), d_off.data_ptr(), peer.target.data_ptr(), sync=False)
                    peer.finish()
                stream.synchronize()
                if rank == 0:
                    same = bool(torch.equal(peer.local, ref))
                    ok &= same
                    print(f"{algo_name} {w}x{h} zoom {zoom}: peer mode {mode}: {'bitwise equal to the 1-GPU render' if same else 'MISMATCH'}", flush=True)
                peer.finish()
                stream.synchronize()
                del peer
            d_out = torch.zeros((3, out_h, out_w), dtype=torch.float32, device=dev)
            ctx.render_planes_device(blk, algo, 3, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=False)
            full = gather_bands(d_out[:, rb:re, :].permute(1, 0, 2).contiguous(), out_h, rank, world)
            stream.synchronize()
            if rank == 0:
                same = bool(torch.equal(full.permute(1, 0, 2), ref))
                ok &= same
                print(f"{algo_name} {w}x{h} zoom {zoom}: NCCL gather: {'bitwise equal to the 1-GPU render' if same else 'MISMATCH'}", flush=True)
        # host-pointer ABI call per rank, output = one page-locked host image shared by all ranks
        shared = SharedHostImage.create((3, out_h, out_w), rank, world, dev)
        if shared is None:
            if rank == 0:
                print(f"{algo_name} {w}x{h} zoom {zoom}: shared host image: NOT AVAILABLE", flush=True)
        else:
            shared.array[...] = -1.0
            dist.barrier()
            ctx.render_planes(blk, algo, [lam[c] for c in range(3)], offsets, [shared.array[c] for c in range(3)])
            dist.barrier()
            if rank == 0:
                same = bool(np.array_equal(shared.array, ref.cpu().numpy()))
                ok &= same
                print(f"{algo_name} {w}x{h} zoom {zoom}: shared host image: {'bitwise equal to the 1-GPU render' if same else 'MISMATCH'}", flush=True)
            dist.barrier()
            shared.close()
        torch.cuda.synchronize()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
