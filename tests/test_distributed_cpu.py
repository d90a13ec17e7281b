"""CPU test of the N>1 path's host logic (world_size 2, gloo): the spp split of
ataraxia_b200.distributed.frame_partition + a float32 sum all-reduce of the per-rank float4
accumulation buffers reproduces the sequential render (sample counts exactly, radiance up to float
reassociation). The per-rank renders here come from the CPU oracle port standing in for the device;
the same split drives atx_render_frames + atx_allreduce_accum (NCCL) on the GPUs.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ataraxia_b200.distributed import frame_partition
    from oracle.bindings import LIGHT_DTYPE, MATERIAL_DTYPE, SPHERE_DTYPE, OraclePort
    gold = np.load(GOLDEN / "cpu_golden.npz")
    port = OraclePort()
    s = gold["small_spheres"].view(SPHERE_DTYPE)
    m = gold["small_materials"].view(MATERIAL_DTYPE)
    l = gold["small_lights"].view(LIGHT_DTYPE)
    rays, pos = gold["small_rays"], gold["small_campos"]
    share = frame_partition(total, rank, world)
    acc = port.render(s, m, l, pos, rays, share.first, share.count, share.stride, 8, True, threads=1)
    t = torch.from_numpy(acc)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # the collective of the path: one sum of float4 buffers
    counts = torch.tensor([share.count])
    dist.all_reduce(counts)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
        np.save(os.path.join(out_dir, "counts.npy"), counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8])
def test_spp_split_allreduce_world2(built, port, tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    assert int(np.load(tmp_path / "counts.npy")[0]) == total
    from oracle.bindings import LIGHT_DTYPE, MATERIAL_DTYPE, SPHERE_DTYPE
    gold = np.load(GOLDEN / "cpu_golden.npz")
    seq = port.render(gold["small_spheres"].view(SPHERE_DTYPE), gold["small_materials"].view(MATERIAL_DTYPE),
                      gold["small_lights"].view(LIGHT_DTYPE), gold["small_campos"], gold["small_rays"], 1, total, 1, 8, True)
    assert (reduced[..., 3] == total).all()                                   # exact sample counts
    assert np.allclose(reduced[..., :3], seq[..., :3], rtol=1e-5, atol=1e-6)  # reassociation only
    rgba_a = port.pack_rgba8(reduced, total)
    rgba_b = port.pack_rgba8(seq, total)
    diff = np.abs(((rgba_a >> 8) & 0xFF).astype(int) - ((rgba_b >> 8) & 0xFF).astype(int))
    assert diff.max() <= 1                                                    # at most 1 LSB in the display image


def test_tile_partition_covers_every_pixel_once():
    """Image-tile split (atx_render_tiles): 8x4 tiles dealt round-robin; every pixel has exactly one owner, the shares
    are balanced to within one tile, and tile_owner agrees with tile_mask."""
    import numpy as np
    from ataraxia_b200.distributed import tile_mask, tile_owner
    for W, H in ((1920, 1080), (161, 91), (8, 4), (7, 3)):
        for world in (1, 2, 3, 8):
            masks = [tile_mask(W, H, world, r) for r in range(world)]
            assert (sum(m.astype(int) for m in masks) == 1).all()
            tiles = [len({(y // 4, x // 8) for y, x in zip(*np.nonzero(m))}) for m in masks]
            assert max(tiles) - min(tiles) <= 1
            for (x, y) in ((0, 0), (W - 1, H - 1), (W // 2, H // 3)):
                assert masks[tile_owner(x, y, W, world)][y, x]
