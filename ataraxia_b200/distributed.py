"""Multi-GPU spp split: every (pixel, frameIndex) sample is independent and the RNG
is a pure function of them (Renderer.cu:300-306), so rank r of R renders frame indices
r+1, r+1+R, ... for the whole image into its own zeroed float4 buffer and the buffers are summed
once (ncclAllReduce float32 sum over NVLink, atx_allreduce_accum). No other data-path collective.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class FrameShare:
    first: int   # first frameIndex of this rank (frame indices start at 1)
    count: int   # how many frames this rank renders
    stride: int  # distance between consecutive frames of this rank


def frame_partition(total_frames: int, rank: int, world: int, first_frame: int = 1) -> FrameShare:
    """Interleaved split: rank r gets first_frame + r + j*world. Balanced to within one frame."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if total_frames < 0:
        raise ValueError("negative frame count")
    count = (total_frames - rank + world - 1) // world if total_frames > rank else 0
    return FrameShare(first_frame + rank, count, world)


def render_split(renderer, total_frames: int, rank: int, world: int, reduce: bool = True) -> FrameShare:
    """Render this rank's share into a zeroed buffer and (optionally) all-reduce it in place."""
    share = frame_partition(total_frames, rank, world)
    renderer.renderFrames(share.first, share.count, share.stride, zero_first=True)
    if reduce and world > 1:
        renderer.allreduceAccum()
    return share


def tile_owner(x: int, y: int, width: int, world: int) -> int:
    """Image-tile split (atx_render_tiles): the rank that renders pixel (x, y). Tiles are 8x4 pixels, numbered row-major,
    and dealt to the ranks round-robin."""
    tiles_x = (width + 7) // 8
    return ((y // 4) * tiles_x + (x // 8)) % world


def tile_mask(width: int, height: int, world: int, rank: int):
    """Boolean (height, width) mask of the pixels rank `rank` of `world` renders under the image-tile split."""
    import numpy as np
    ys, xs = np.mgrid[0:height, 0:width]
    return ((ys // 4) * ((width + 7) // 8) + (xs // 8)) % world == rank


def render_tiles(renderer, first_frame: int, n_frames: int, zero_first: bool = True) -> None:
    """Collective: every rank renders all n_frames of its tiles and stores them into every rank's image (NVLink peer stores)."""
    renderer.renderTiles(first_frame, n_frames, zero_first)
