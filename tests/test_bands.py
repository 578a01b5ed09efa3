"""Host logic of the multi-rank split: strip partition, padding, unstripe, and a world-size-2
gloo gather where each rank 'renders' its strips with the CPU oracle."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden_frame, load_golden_scene
from gp1_raytracer_2223_b200 import _lib, bands, build
from oracle import rt_oracle

build.ensure()          # rt_host_arrive_and_wait lives in the C-ABI library (it needs no GPU)


@pytest.mark.parametrize("height", [1, 7, 8, 9, 77, 480, 2160])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_every_row_once(height, world):
    seen = np.zeros(height, dtype=int)
    for rank in range(world):
        rows = bands.rows_of_rank(height, world, rank)
        assert all(b < bands.band_rows(height, world) for _, b in rows)
        assert len({b for _, b in rows}) == len(rows)
        for y, _ in rows:
            seen[y] += 1
    assert (seen == 1).all()


def _render_band(scene, width, height, world, rank):
    band = np.zeros((bands.band_rows(height, world), width), dtype=np.uint32)
    for local, strip in enumerate(bands.strips_of_rank(height, world, rank)):
        y0 = strip * bands.STRIP_ROWS
        n = min(bands.STRIP_ROWS, height - y0)
        band[local * bands.STRIP_ROWS: local * bands.STRIP_ROWS + n] = rt_oracle.render(scene, width, height, row_begin=y0, row_count=n)
    return band


@pytest.mark.parametrize("world", [2, 3, 8])
def test_unstripe_reassembles_reference_frame(world):
    scene = load_golden_scene("bunny_333x77")
    want = load_golden_frame("bunny_333x77")
    stacked = np.stack([_render_band(scene, 333, 77, world, r) for r in range(world)])
    assert np.array_equal(bands.unstripe_numpy(stacked, 333, 77, world), want)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = load_golden_scene("w4ref_101x203")
        band = torch.from_numpy(_render_band(scene, 101, 203, world, rank).view(np.int32))
        gathered = bands.gather_bands(band, rank, world)
        if rank == 0:
            frame = bands.unstripe_numpy(gathered.numpy().view(np.uint32), 101, 203, world)
            q.put(frame)
        else:
            assert gathered is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_matches_reference():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    frame = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(frame, load_golden_frame("w4ref_101x203"))


def _surface_worker(rank, world, port, path, frames, q):
    """One rank of the direct present, the CPU oracle standing in for the GPU: its strips go straight into the shared
    surface, completion by the arrival words (gloo only carries the rendezvous and the surface's name)."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        name = [path if rank == 0 else None]
        if rank == 0:
            surface = bands.SharedSurface(101, 203, world, rank, path, create=True)
        dist.broadcast_object_list(name, src=0)
        if rank != 0:
            surface = bands.SharedSurface(101, 203, world, rank, name[0], create=False)
        dist.barrier()
        if rank == 0:
            surface.unlink()
        scene = load_golden_scene("w4ref_101x203")
        for k in range(frames):
            for strip in bands.strips_of_rank(203, world, rank):
                y0 = strip * bands.STRIP_ROWS
                n = min(bands.STRIP_ROWS, 203 - y0)
                surface.frame[y0:y0 + n] = rt_oracle.render(scene, 101, 203, row_begin=y0, row_count=n)
            # odd frames through the interpreted spin, even ones through the C ABI's rt_host_arrive_and_wait
            assert surface.arrive_and_wait(lib=_lib.load() if k % 2 == 0 else None) == k + 1
            if rank == 0:
                q.put(surface.frame.copy())
            dist.barrier()                              # the test's own pacing: rank 0 has copied before anyone overwrites
    finally:
        dist.destroy_process_group()


def test_gloo_world2_direct_present_into_shared_surface(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path)
    path = os.path.join(shm_dir, f"rt_b200_test_surface_{os.getpid()}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    frames = 2
    procs = [ctx.Process(target=_surface_worker, args=(r, 2, port, path, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(frames)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert not os.path.exists(path)
    want = load_golden_frame("w4ref_101x203")
    for frame in got:
        assert np.array_equal(frame, want)


def test_shared_surface_handshake_times_out_instead_of_hanging(tmp_path):
    path = str(tmp_path / "surface")
    s = bands.SharedSurface(16, 16, 2, 0, path, create=True)
    with pytest.raises(TimeoutError):
        s.arrive_and_wait(timeout_s=0.05)            # rank 1 never arrives
    with pytest.raises(TimeoutError):
        s.arrive_and_wait(timeout_s=0.05, lib=_lib.load())
