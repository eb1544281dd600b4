"""Frame partitioning across the GPUs of one box (new capability — the reference is single-GPU).

Primary scheme: SAMPLE SHARDING. Every rank holds the full scene + BVH (tens of MB for Sponza-class scenes, rebuilt locally in
a few ms) and renders its own frames with a disjoint slice of the reference's frameCount sequence 1, 3, 5, ... :
rank g renders frameCount = 1 + 2*(g + k*G), k = 0, 1, ... Each rank accumulates its samples in fp32 (a sum, not a running mean)
and ONE collective — a sum-reduce of the W*H*4 float accumulation buffer — produces the image. There is no other exchange.

Alternative for single-frame latency: IMAGE BANDS with a 60-pixel halo (ReSTIR spatial radius 30 x 2 iterations), rendered
redundantly so no halo exchange is needed; `band_partition` computes the bands, the gather is one collective at the end.

torch.distributed is the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests); nothing here touches pixels except
through the reduce.
"""
from __future__ import annotations

from dataclasses import replace
from typing import List, Tuple

RESTIR_HALO = 60          # spatial radius 30 px x 2 iterations (ReSTIRData.h:49,56)


def shard_settings(settings, rank: int, world: int):
    """Settings of rank `rank`: blend (accumulate) mode and a disjoint frameCount stream."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return replace(settings, blend_output=True, first_frame_count=2 * rank, frame_count_stride=2 * world)


def frame_counts(rank: int, world: int, frames: int) -> List[int]:
    """The reference's frameCount values rank `rank` consumes (the counter advances twice per frame: 1, 3, 5, ...)."""
    return [2 * rank + 1 + 2 * world * k for k in range(frames)]


def split_frames(total_frames: int, world: int) -> List[int]:
    """Frames per rank for a fixed total sample count (strong scaling of a progressive render, config C5)."""
    base, rem = divmod(total_frames, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def band_partition(height: int, world: int, halo: int = RESTIR_HALO) -> List[Tuple[int, int, int, int]]:
    """Row bands (y0, y1) owned by each rank and the rows (h0, h1) it must render so that ReSTIR reuse inside its band never
    reads a pixel it did not compute."""
    out = []
    for r in range(world):
        y0, y1 = height * r // world, height * (r + 1) // world
        out.append((y0, y1, max(0, y0 - halo), min(height, y1 + halo)))
    return out


def reduce_accumulation(accum, dst: int = 0, group=None):
    """The one collective of the path: sum the per-rank fp32 accumulation buffers onto rank `dst` (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum
