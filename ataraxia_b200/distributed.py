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
