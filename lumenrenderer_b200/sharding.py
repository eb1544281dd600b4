"""Frame partitioning across the GPUs of one box (new capability — the reference is single-GPU).

Primary scheme: SAMPLE SHARDING. Every rank holds the full scene + BVH (tens of MB for Sponza-class scenes, rebuilt locally in
a few ms) and renders its own frames with a disjoint slice of the reference's frameCount sequence 1, 3, 5, ... :
rank g renders frameCount = 1 + 2*(g + k*G), k = 0, 1, ... Each rank accumulates its samples in fp32 (a sum, not a running mean)
and ONE collective — a sum-reduce of the W*H*4 float accumulation buffer — produces the image. There is no other exchange.

Alternative for single-frame latency: IMAGE BANDS with a 60-pixel halo (ReSTIR spatial radius 30 x 2 iterations), rendered
redundantly so no halo exchange is needed: `band_settings` turns a renderer into the producer of one row band of the full frame
(LbSettings::band_row0 / band_full_height), `gather_bands` collects the owned rows on one rank at the end of the frame.

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


def band_partition(height: int, world: int, halo: int = RESTIR_HALO, width: int = 0) -> List[Tuple[int, int, int, int]]:
    """Row bands (y0, y1) owned by each rank and the rows (h0, h1) it must render so that ReSTIR reuse inside its band never
    reads a pixel it did not compute. With `width` given, h0 is lowered until h0 * width is a multiple of 256 (the RIS light-bag
    group is 256 consecutive pixels of the FULL frame, LbSettings::band_row0)."""
    from math import gcd
    step = 256 // gcd(width, 256) if width else 1
    out = []
    for r in range(world):
        y0, y1 = height * r // world, height * (r + 1) // world
        h0 = max(0, y0 - halo)
        out.append((y0, y1, h0 - h0 % step, min(height, y1 + halo)))
    return out


def band_settings(settings, rank: int, world: int, halo: int = RESTIR_HALO):
    """Settings of the renderer that produces rank `rank`'s band (+ halo) of the frame described by `settings`, and the band
    (y0, y1, h0, h1). Camera, jitter, motion vectors and random streams stay keyed on full-frame pixel positions, so the pixels of
    rows y0..y1 are the ones a single renderer would produce (bit for bit while the ReSTIR history they depend on lies inside the halo:
    the first two frames after a history reset; afterwards the outer `halo` rows of a band reuse a slightly different neighbourhood).
    The halo rows (outside band_own_*) only produce what a ReSTIR neighbour needs of them — primary surface record and reservoirs — and
    spawn no NEE or bounce rays. Same arithmetic as lb_band_settings of the C ABI (csrc/lb_multigpu.cpp)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    y0, y1, h0, h1 = band_partition(settings.height, world, halo, settings.width)[rank]
    return replace(settings, height=h1 - h0, band_row0=h0, band_full_height=settings.height, band_own_row0=y0, band_own_rows=y1 - y0), (y0, y1, h0, h1)


def gather_bands(band_rows, full_frame, bands, rank: int, dst: int = 0, group=None):
    """The one collective of the band scheme: every rank sends the rows it owns to `dst`. `band_rows` is this rank's rendered frame
    as an (h1-h0, W, 4) tensor, `full_frame` the (H, W, 4) tensor on `dst` (ignored elsewhere), `bands` = band_partition(...).
    Point-to-point sends, because the bands differ in size when H is not a multiple of the world size."""
    import torch.distributed as dist
    y0, y1, h0, _ = bands[rank]
    own = band_rows[y0 - h0:y1 - h0]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        full_frame[y0:y1].copy_(own)
        return full_frame
    ops = []
    if rank == dst:
        full_frame[y0:y1].copy_(own)
        for r, (a, b, _, _) in enumerate(bands):
            if r != dst:
                ops.append(dist.P2POp(dist.irecv, full_frame[a:b], r, group))
    else:
        ops.append(dist.P2POp(dist.isend, own.contiguous(), dst, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return full_frame


def reduce_accumulation(accum, dst: int = 0, group=None):
    """The one collective of the path: sum the per-rank fp32 accumulation buffers onto rank `dst` (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum
