"""Row-strip partition of a frame across ranks and the gather to rank 0.

SURVEY.md 8(e): every pixel is independent, the scene is replicated, the only exchange is
the gather of the finished rows on rank 0.  The frame is cut into strips of ``STRIP_ROWS``
rows (one CTA row of the kernel) dealt round-robin: rank r renders strips r, r + world, ...
into a packed band, which balances the expensive lower-middle rows of the bunny scene over
all ranks.  Bands are padded to ``strips_per_rank`` strips so the gather is uniform.

Everything here is index arithmetic plus torch.distributed calls (NCCL on GPUs, gloo in the
CPU tests); the only device work besides the collective is the library's unstripe kernel.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

STRIP_ROWS = 8


def total_strips(height: int) -> int:
    return (height + STRIP_ROWS - 1) // STRIP_ROWS


def strips_per_rank(height: int, world: int) -> int:
    return (total_strips(height) + world - 1) // world


def strips_of_rank(height: int, world: int, rank: int):
    return list(range(rank, total_strips(height), world))


def band_rows(height: int, world: int) -> int:
    """Rows in one (padded) packed band."""
    return strips_per_rank(height, world) * STRIP_ROWS


def rows_of_rank(height: int, world: int, rank: int):
    """(frame_row, band_row) pairs this rank owns, in band order."""
    out = []
    for local, strip in enumerate(strips_of_rank(height, world, rank)):
        for r in range(STRIP_ROWS):
            y = strip * STRIP_ROWS + r
            if y < height:
                out.append((y, local * STRIP_ROWS + r))
    return out


def unstripe_numpy(bands: np.ndarray, width: int, height: int, world: int) -> np.ndarray:
    """Reference (host) statement of the unstripe step: bands is (world, band_rows, width)."""
    frame = np.empty((height, width), dtype=bands.dtype)
    for rank in range(world):
        for y, b in rows_of_rank(height, world, rank):
            frame[y] = bands[rank, b]
    return frame


def gather_bands(band, rank: int, world: int, dst: int = 0, group=None):
    """torch.distributed gather of equally sized bands on ``dst``; returns (world, ...) on dst, None elsewhere."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return band.unsqueeze(0)
    if rank == dst:
        out = torch.empty((world,) + tuple(band.shape), dtype=band.dtype, device=band.device)
        dist.gather(band, list(out.unbind(0)), dst=dst, group=group)
        return out
    dist.gather(band, None, dst=dst, group=group)
    return None


class SharedSurface:
    """The host surface of a one-process-per-GPU launch: one shared-memory mapping that rank 0 owns (like the window
    surface of the reference's process, source/Renderer.cpp:27-29) and every rank opens, so that each rank can copy the
    strips it rendered straight into it (rt_render_strips_to_host) - N PCIe links instead of GPU 0's one.

    Next to the pixels sits one arrival word per rank (own cache line, written only by that rank): rank r stores k
    once its strips of frame k are in the surface; the frame is complete when every word has reached k.  That is the
    whole exchange of this leg: host memory, no collective.
    """

    def __init__(self, width: int, height: int, world: int, rank: int, path: str, create: bool):
        import mmap
        self.width, self.height, self.world, self.rank, self.path = width, height, world, rank, path
        self.surface_bytes = width * height * 4
        total = self.surface_bytes + 64 * world
        if create:
            with open(path, "wb") as f:
                f.truncate(total)
        self._file = open(path, "r+b")
        self._map = mmap.mmap(self._file.fileno(), total)
        self.frame = np.frombuffer(self._map, dtype=np.uint32, count=width * height).reshape(height, width)
        self.arrived = np.frombuffer(self._map, dtype=np.int64, offset=self.surface_bytes, count=8 * world)[::8]
        self.presented = 0

    @property
    def ptr(self) -> int:
        return self.frame.ctypes.data

    @property
    def pitch_bytes(self) -> int:
        return 4 * self.width

    def unlink(self) -> None:
        """Rank 0, once every rank has opened the mapping: the name goes, the mappings keep the memory alive."""
        import os
        os.unlink(self.path)

    def arrive_and_wait(self, timeout_s: float = 60.0, lib=None) -> int:
        """This rank's strips of the next frame are in the surface; returns when everybody's are.  With `lib` (the
        loaded C ABI) the spin runs in rt_host_arrive_and_wait instead of the interpreter."""
        import time
        self.presented += 1
        if lib is not None:
            rc = lib.rt_host_arrive_and_wait(self.arrived.ctypes.data, 8, self.rank, self.world, self.presented, timeout_s)
            if rc != 0:
                raise TimeoutError(f"rank {self.rank}: frame {self.presented} incomplete, arrivals {self.arrived.tolist()}")
            return self.presented
        self.arrived[self.rank] = self.presented
        deadline = time.monotonic() + timeout_s
        spins = 0
        while int(self.arrived.min()) < self.presented:
            spins += 1
            if (spins & 0xFFF) == 0 and time.monotonic() > deadline:
                raise TimeoutError(f"rank {self.rank}: frame {self.presented} incomplete, arrivals {self.arrived.tolist()}")
        return self.presented
