"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun). Every rank renders its share of
the frames into a zeroed buffer, the float4 buffers are summed with atx_allreduce_accum (NCCL over
NVLink) and rank 0 compares with the same frames rendered sequentially on its own GPU."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import ataraxia_b200 as atx  # noqa: E402
from ataraxia_b200.distributed import render_split  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total, W, H, bounces = 2 * world + 3, 320, 180, 8
    scene = atx.Utils.importScene(str(ROOT / "tests" / "golden" / "sample_scene.json"))
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(local)
    r.setSettings(atx.Settings(True, False, bounces))
    r.onResize(W, H); cam.Resize(W, H)
    r.uploadScene(scene); r.setCamera(cam)
    uid = [atx.Renderer.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    r.commInitRank(world, rank, uid[0])
    # progressive preview: every rank renders the first half of its share, the out-of-place sum is looked at, the
    # ranks' own buffers are untouched, and rendering goes on to the full share
    from ataraxia_b200.distributed import frame_partition
    sh = frame_partition(total, rank, world)
    half = sh.count // 2
    r.renderFrames(sh.first, half, sh.stride, zero_first=True)
    own_before = r.getAccumulation()
    r.allreducePreview()
    done = torch.tensor([half], device=f"cuda:{local}"); dist.all_reduce(done)
    prev_acc, prev_rgba = r.getPreview(int(done.item()))
    preview_ok = bool((prev_acc[..., 3] == int(done.item())).all()) and bool((r.getAccumulation().view(np.uint32) == own_before.view(np.uint32)).all())
    preview_ok &= bool(((prev_rgba >> 24) == 255).all())
    r.renderFrames(sh.first + half * sh.stride, sh.count - half, sh.stride, zero_first=False)
    continued = r.getAccumulation()
    share = render_split(r, total, rank, world)          # frames rank+1, rank+1+world, ... then the sum across ranks
    reduced = r.getAccumulation()
    rgba = r.getRGBA8(divisor=total)
    ok = preview_ok
    # default transport: one kernel over NVLink peer memory (separate processes on one NVSwitch domain can always map each other)
    p2p_ok = r.lastReduceKind() == atx.REDUCE_PEER_MEMORY
    # the same sum through ncclAllReduce: equal up to the order of N float additions; and again through peer memory
    # (second epoch of the flag barriers): bit-identical to the first time
    r.setTuning(atx.TUNE_REDUCE, 1)
    render_split(r, total, rank, world)
    via_nccl = r.getAccumulation()
    p2p_ok &= r.lastReduceKind() == atx.REDUCE_NCCL
    p2p_ok &= bool((via_nccl[..., 3] == reduced[..., 3]).all()) and bool(np.allclose(via_nccl[..., :3], reduced[..., :3], rtol=1e-6, atol=1e-7))
    r.setTuning(atx.TUNE_REDUCE, 0)
    for _ in range(3):
        render_split(r, total, rank, world)
    p2p_ok &= bool((r.getAccumulation().view(np.uint32) == reduced.view(np.uint32)).all()) and r.lastReduceKind() == atx.REDUCE_PEER_MEMORY
    # a resize replaces the accumulation buffer: the peer mappings are set up again at the next reduce
    r.onResize(W + 32, H); cam.Resize(W + 32, H); r.setCamera(cam)
    render_split(r, total, rank, world)
    wide = r.getAccumulation()
    p2p_ok &= bool((wide[..., 3] == total).all()) and r.lastReduceKind() == atx.REDUCE_PEER_MEMORY
    r.onResize(W, H); cam.Resize(W, H); r.setCamera(cam)
    render_split(r, total, rank, world)
    p2p_ok &= bool((r.getAccumulation().view(np.uint32) == reduced.view(np.uint32)).all())
    # per-rank checkpoint of an spp-split render: half of the share, save, load, rest of the share == the whole share
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        r.renderFrames(sh.first, half, sh.stride, zero_first=True)
        path = f"{td}/rank{rank}.ckpt"
        r.saveCheckpoint(path, next_frame_index=sh.first + half * sh.stride, frame_stride=sh.stride)
        r.renderFrames(1, 1, 1, zero_first=True)             # clobber
        nxt, stride = r.loadCheckpoint(path)
        r.renderFrames(nxt, sh.count - half, stride, zero_first=False)
        p2p_ok &= bool((r.getAccumulation().view(np.uint32) == continued.view(np.uint32)).all())
    # image-tile split: every rank renders ALL frames of its interleaved 8x4 tiles and stores them into every rank's image over
    # NVLink while rendering; the image is bit-identical to a single-GPU render, on every rank, also when continuing
    r.renderTiles(1, total, zero_first=True)
    tiles1 = r.getAccumulation()
    tile_kind = r.lastReduceKind()
    r.renderTiles(total + 1, 3, zero_first=False)
    tiles2 = r.getAccumulation()
    r.renderFrames(1, total, 1, zero_first=True)
    p2p_ok &= bool((r.getAccumulation().view(np.uint32) == tiles1.view(np.uint32)).all()) and tile_kind == atx.REDUCE_PEER_MEMORY
    r.renderFrames(total + 1, 3, 1, zero_first=False)
    p2p_ok &= bool((r.getAccumulation().view(np.uint32) == tiles2.view(np.uint32)).all())
    r.setTuning(atx.TUNE_REDUCE, 1)                         # without peer mappings: own tiles kept, the rest zeroed, ncclAllReduce
    r.renderTiles(1, total, zero_first=True)
    r.renderTiles(total + 1, 3, zero_first=False)
    p2p_ok &= bool((r.getAccumulation().view(np.uint32) == tiles2.view(np.uint32)).all()) and r.lastReduceKind() == atx.REDUCE_NCCL
    r.setTuning(atx.TUNE_REDUCE, 0)
    ok &= p2p_ok
    if not p2p_ok:
        print(f"rank {rank}: peer-memory reduce checks failed (last kind {r.lastReduceKind()})", flush=True)
    # the share rendered in two launches around the preview equals the share rendered in one (render_split re-renders it)
    r.renderFrames(sh.first, sh.count, sh.stride, zero_first=True)
    ok &= bool((r.getAccumulation().view(np.uint32) == continued.view(np.uint32)).all())
    if rank == 0:
        r.renderFrames(1, total, 1, zero_first=True)
        seq = r.getAccumulation()
        rgba_seq = r.getRGBA8(divisor=total)
        ok &= bool((reduced[..., 3] == total).all())                                    # sample counts exact
        ok &= bool(np.allclose(reduced[..., :3], seq[..., :3], rtol=2e-6, atol=1e-6))   # float reassociation only
        d = np.abs(((rgba >> 8) & 0xFF).astype(int) - ((rgba_seq >> 8) & 0xFF).astype(int))
        ok &= bool(d.max() <= 1)
        print(f"MGPU world={world} total={total} share0={share.count} preview_ok={preview_ok} p2p_ok={p2p_ok} counts_ok={(reduced[..., 3] == total).all()} "
              f"max_abs_diff={np.abs(reduced[..., :3] - seq[..., :3]).max():.3e} rgba_lsb={d.max()} ok={ok}", flush=True)
    # every rank holds the same reduced buffer
    t = torch.from_numpy(reduced.copy()).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all().item())
    r.commDestroy(); r.close()
    dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
