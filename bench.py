#!/usr/bin/env python
"""Benchmark of the hot path: Renderer::Render of Scene_W4_BunnyScene at 3840x2160,
Combined lighting, shadows on (BASELINE.json configs[4], the north-star target).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one frame.  Metric: Mrays/s (primary + shadow rays of the frame / time).
  value      device-timed (CUDA events on the launching stream), scene resident in HBM, frame left in
             HBM on rank 0.  N > 1: every rank renders its 8-row strips straight into rank 0's frame
             (CUDA-IPC mapped peer memory, 128-bit stores over NVLink: render and gather are one kernel),
             and bumps a signal word behind the frame; rank 0's stream waits on that word
             (cuStreamWaitValue32): no collective on the data path.  RT_BENCH_GATHER=fused-barrier uses a
             one-element NCCL all-reduce as the barrier instead, RT_BENCH_GATHER=nccl the unfused form
             (packed bands + NCCL gather + unstripe kernel).
             Headline = the library's default mesh path: the reference's shipped BVH walk over the
             reference's own nodes (source/Utils.h:246-297).  The slab + every-triangle body the north
             star names (source/Utils.h:298-325) is measured the same way and reported under
             "north_star_slab_linear" with its own roofline.
  e2e        the same frame through the reference-facing C ABI with HOST buffers: per step the
             re-transformed mesh goes host -> device (rt_upload_mesh, what Scene::Update produces each
             frame) and the finished frame comes device -> host (rt_render into pinned memory)
  roofline   FP32: algorithmic FLOP of the frame (SURVEY.md 8(d) table x the counters build's event
             counts) / kernel time, against the FP32 issue peak measured live on the same GPU
  cpu_baseline  the reference's own parallel CPU loop (oracle/_ref/ref_render, the unmodified
             reference sources compiled in the build container) on this box's host cores

`--impl reference` prints the reference arm: the same metric from the reference binary alone.
The oracle is only ever the checker / the reported baseline here, never the measured product.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "Scene_W4_BunnyScene 3840x2160 Combined lighting, shadows on, pose after Initialize()"
FIXTURE = os.path.join(ROOT, "tests", "golden", "bunny_4k.rtsc")
WIDTH, HEIGHT = 3840, 2160
RAYS_PER_FRAME = WIDTH * HEIGHT * 4          # 1 primary + 3 shadow rays per pixel (every pixel hits; re-checked below)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_render")
L2_FLUSH_BYTES = 256 << 20
KERNEL_NOTE = "persistent warps (one pixel per thread, 8x4 warp tiles pulled off a device queue; 128-thread CTAs, 8 per SM, 64 registers); rt_render overlaps the present copies with rendering (band watcher in the kernel, copies issued by the host thread as bands complete)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def merge(self, other):
        self.samples += other.samples
        self.reasons |= other.reasons

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference compiled under oracle/_ref
# ------------------------------------------------------------------------------------------------
def run_reference(frames: int, warmup: int, threads: int = 0):
    """Times Renderer::Render of the reference binary on the host cores.  Returns the driver's JSON."""
    args = [REF_BIN, "--scene", "W4_Bunny", "--width", str(WIDTH), "--height", str(HEIGHT), "--mode", "3",
            "--shadows", "1", "--frames", str(frames), "--warmup", str(warmup)]
    # all host cores unless told otherwise: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently turn the reference's parallel loop into a serial one
    args += ["--threads", str(threads or host_cores())]
    env = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
    out = subprocess.run(args, check=True, capture_output=True, text=True, env=env).stdout
    return json.loads(out.strip().splitlines()[-1])


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_port(frames: int, warmup: int):
    """Fallback when oracle/_ref is absent: the C restatement (kind 'port')."""
    from gp1_raytracer_2223_b200 import load_rtsc
    from oracle import rt_oracle
    scene = load_rtsc(FIXTURE)
    ms = []
    for i in range(warmup + frames):
        t0 = time.perf_counter()
        rt_oracle.render(scene, WIDTH, HEIGHT, threads=host_cores())
        if i >= warmup:
            ms.append((time.perf_counter() - t0) * 1e3)
    return {"ms": ms, "ms_median": statistics.median(ms), "threads": host_cores(),
            "path": "C restatement, slab + linear triangle loop"}


def reference_available():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def bounded_reference_run(steps: int, warmup: int, budget_s: float):
    """Runs at most `steps` frames, fewer if they would not fit in `budget_s` seconds."""
    runner = run_reference if reference_available() else run_port
    probe = runner(1, 1)                       # 1 warm-up + 1 timed frame: estimate the frame time
    est_ms = probe["ms_median"]
    frames = max(1, min(steps, int(budget_s * 1e3 / max(est_ms, 1e-3))))
    res = runner(frames, min(warmup, 2))
    kind = "reference" if reference_available() else "port"
    return res, frames, kind


def run_reference_scene(scene: str, width: int, height: int, frames: int, warmup: int, extra=()):
    """Any reference scene through the compiled reference (oracle/_ref/ref_render), all host cores."""
    args = [REF_BIN, "--scene", scene, "--width", str(width), "--height", str(height), "--mode", "3", "--shadows", "1",
            "--frames", str(frames), "--warmup", str(warmup), "--threads", str(host_cores())] + list(extra)
    env = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
    out = subprocess.run(args, check=True, capture_output=True, text=True, env=env).stdout
    return json.loads(out.strip().splitlines()[-1])


DATA = "reference scene fixture tests/golden/bunny_4k.rtsc (dumped from the reference's Scene_W4_BunnyScene::Initialize; deterministic, no RNG)"


def bench_config():
    """The workload both arms run, word for word the same dict in both lines (the driver compares them)."""
    return {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "rays_per_frame": RAYS_PER_FRAME, "triangles": 292, "lights": 3,
            "lighting_mode": "Combined", "shadows": True, "mesh_path": "BVH (what the reference ships, source/Utils.h:296-297)",
            "l2": f"GPU arm: {L2_FLUSH_BYTES >> 20} MiB memset between timed steps (outside the event pairs); the CPU arm has no device cache to flush"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    warmup = max(args.warmup, 3)              # same warm-up rule as the GPU arm
    runner = run_reference if reference_available() else run_port
    kind = "reference" if reference_available() else "port"
    est_ms = runner(1, 1)["ms_median"]      # 1 warm-up + 1 timed frame: how long is a frame on this box?
    frames = max(1, min(args.steps, int(150e3 / max(est_ms, 1e-3)) - warmup))
    res = runner(frames, warmup)
    ms = statistics.mean(res["ms"])
    value = RAYS_PER_FRAME / (ms * 1e-3) / 1e6
    sample = f"{frames} full {WIDTH}x{HEIGHT} frames of Renderer::Render" + ("" if frames == args.steps else f" (of {args.steps} requested steps; bounded to ~150 s)")
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": frames, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": DATA,
        "config": bench_config(),
        "step_ms": {"min": min(res["ms"]), "median": statistics.median(res["ms"]), "max": max(res["ms"])},
        "reference_path": res.get("path"),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": res["threads"], "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout: keep the real stdout aside and send everything else
    # (NCCL banners, library chatter) to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        sys.exit("bench.py measures the CUDA path; no GPU is visible (there is no CPU fallback)")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs a torchrun launch with {args.gpus} ranks (one process per GPU)")
        sys.exit(f"--gpus {args.gpus} does not match WORLD_SIZE {world}")
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host-side barriers (no kernel spinning on anybody's GPU) for the legs rank 0 runs alone on ALL devices
        host_group = dist.new_group(backend="gloo")

    from gp1_raytracer_2223_b200 import Renderer, bands, build, load_rtsc
    from gp1_raytracer_2223_b200.flops import algorithmic_flops, rays
    build.ensure()                              # rebuilds when the library is not what the sources in the tree compile to
    library = build.provenance()

    scene = load_rtsc(FIXTURE)
    mesh = scene.meshes[0]
    r = Renderer(WIDTH, HEIGHT, device_ids=[local_rank])
    r.SetScene(scene)
    # A real (non-default) stream: the C ABI treats a NULL stream as "use the context's own stream and block",
    # and the timing events must sit on the stream the kernels are launched on.
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload accounting from the counters build (rank 0 only; not timed) -----------------
    PATHS = {"bvh": 2, "slab_linear": 1}
    flop_per_frame = {}
    if rank == 0:
        for pname, pid in PATHS.items():
            counters = r.count_frame(mesh_path=pid)
            flop_per_frame[pname] = algorithmic_flops(counters, 3)
            assert rays(counters) == RAYS_PER_FRAME, (rays(counters), RAYS_PER_FRAME)
    gather_mode = os.environ.get("RT_BENCH_GATHER", "fused")      # fused | fused-barrier | nccl
    fused = world > 1 and gather_mode != "nccl"
    signals = fused and gather_mode == "fused"
    expected = [0]
    present_no = [0]
    PRESENT_BANDS = 16
    spr = bands.strips_per_rank(HEIGHT, world)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    host_frame = torch.empty((HEIGHT, WIDTH), dtype=torch.int32).pin_memory() if rank == 0 else None
    # N > 1, end to end: the host surface is one shared-memory mapping (rank 0 owns it, like the window surface of the
    # reference's process); every rank copies the strips it rendered into it over its own PCIe link
    # (rt_render_strips_to_host).  RT_BENCH_PRESENT=gather: all strips to GPU 0 over NVLink, GPU 0 presents (one link).
    present_mode = os.environ.get("RT_BENCH_PRESENT", "direct") if world > 1 else "single"
    shared_surface = None
    if present_mode == "direct":
        # any rank that cannot create / open / map the surface sends everybody back to the gather flow together
        import tempfile
        name, surface = [None], None
        if rank == 0:
            for folder in ("/dev/shm", tempfile.gettempdir()):
                try:
                    candidate = os.path.join(folder, f"rt_b200_surface_{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}")
                    surface = bands.SharedSurface(WIDTH, HEIGHT, world, rank, candidate, create=True)
                    name = [candidate]
                    break
                except OSError as exc:
                    print(f"rank 0: no shared surface in {folder}: {exc}", file=sys.stderr)
        dist.broadcast_object_list(name, src=0)
        opened = torch.ones(1, dtype=torch.int32, device="cuda")
        if name[0] is None:
            opened[0] = 0
        elif rank != 0:
            try:
                surface = bands.SharedSurface(WIDTH, HEIGHT, world, rank, name[0], create=False)
            except OSError as exc:
                print(f"rank {rank}: cannot open the shared surface: {exc}", file=sys.stderr)
                opened[0] = 0
        dist.all_reduce(opened, op=dist.ReduceOp.MIN)
        everybody_opened = int(opened) == 1         # reading the value waits for the collective: every rank has tried by now
        if rank == 0 and name[0] is not None:
            surface.unlink()                        # the mappings keep it alive; nothing is left behind
        if everybody_opened:
            shared_surface = surface.frame
            # every rank pins ITS mapping of the surface for direct copies; the mapping outlives the context (rt_b200.h)
            r.register_surface(surface.ptr, surface.surface_bytes)
        else:
            present_mode, surface = "gather", None
    band = frame_dev = gathered = None
    frame_ptr = 0
    token = torch.zeros(1, dtype=torch.int32, device="cuda")
    if world == 1:
        frame_dev = torch.empty((HEIGHT, WIDTH), dtype=torch.int32, device="cuda")
    if world > 1 and fused:
        # rank 0 owns the frame, the others map it (CUDA IPC).  If any rank cannot (no peer access between
        # the GPUs, IPC disabled in the container), every rank falls back to the NCCL gather together.
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            handle = [r.frame_export() if rank == 0 else None]
        except Exception as exc:                                   # noqa: BLE001
            print(f"rank {rank}: frame export failed: {exc}", file=sys.stderr)
            handle, ok[0] = [None], 0
        dist.broadcast_object_list(handle, src=0)
        if rank != 0 and handle[0] is not None:
            try:
                frame_ptr = r.frame_import(handle[0])
            except Exception as exc:                               # noqa: BLE001
                print(f"rank {rank}: frame import failed: {exc}", file=sys.stderr)
                ok[0] = 0
        elif rank != 0:
            ok[0] = 0
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            if frame_ptr:
                r.frame_release(frame_ptr)
                frame_ptr = 0
            fused = signals = False
    if world > 1 and not fused:
        band = torch.empty((spr * bands.STRIP_ROWS, WIDTH), dtype=torch.int32, device="cuda")
        if rank == 0:
            frame_dev = torch.empty((HEIGHT, WIDTH), dtype=torch.int32, device="cuda")
            gathered = torch.empty((world,) + tuple(band.shape), dtype=torch.int32, device="cuda")

    import lzma
    golden = None
    if rank == 0:
        with open(os.path.join(ROOT, "tests", "golden", "bunny_4k.frame.xz"), "rb") as f:
            planar = np.frombuffer(lzma.decompress(f.read()), dtype=np.uint8).reshape(3, HEIGHT, WIDTH).astype(np.uint32)
        golden = (planar[0] << 16) | (planar[1] << 8) | planar[2]
    POISON = 0xDEADBEEF

    def frame_check_of(produced):
        return {"differing_pixels_vs_reference_frame": int((produced != golden).sum()), "pixels": WIDTH * HEIGHT,
                "poisoned_pixels_left": int((produced == POISON).sum())}

    def poison_device_frame():
        """Before a device-timed leg: the frame buffer that leg fills holds a poison pattern, so the check after the leg
        cannot pass on what an earlier leg left there."""
        barrier()
        if rank == 0:
            if frame_dev is not None:
                frame_dev.fill_(POISON - (1 << 32))
            else:
                r.clear_frame(POISON)
        barrier()

    def device_frame_check():
        """After a device-timed leg: the frame that leg left in rank 0's HBM, against the frame the reference rendered."""
        barrier()
        if rank != 0:
            return None
        if frame_dev is not None:
            produced = frame_dev.cpu().numpy().view(np.uint32)
        else:
            produced = r.download()               # rank 0's exported frame (rt_download_frame)
        return frame_check_of(produced)

    def render_step():
        """This rank's kernel launch of one frame."""
        if world == 1:
            r.render_strips_device(0, 1, frame_dev.data_ptr(), stream)
        elif fused:
            r.render_strips_to_frame(rank, world, frame_ptr, stream)
        else:
            r.render_strips_device(rank, world, band.data_ptr(), stream)

    def exchange_step():
        """What makes the frame complete in rank 0's HBM."""
        if world == 1:
            return
        if fused and signals:
            # completion over the same peer mapping: the other ranks bump the frame's signal word after their
            # strips, rank 0's stream waits until all of them have (no collective on the step)
            if rank == 0:
                expected[0] += world - 1
                r.frame_wait(expected[0], stream)
            else:
                r.frame_signal(frame_ptr, stream)
        elif fused:
            dist.all_reduce(token)                 # barrier, stream-ordered after every rank's kernel
        elif rank == 0:
            dist.gather(band, list(gathered.unbind(0)), dst=0)
            r.unstripe_device(gathered.data_ptr(), frame_dev.data_ptr(), world, spr, stream)
        else:
            dist.gather(band, None, dst=0)

    def device_step():
        """Inputs resident; result = the whole frame in rank 0's HBM."""
        render_step()
        exchange_step()

    e2e_parts = [0.0] * 6      # host seconds in upload / render call / handshake, device seconds kernel / kernel + copies, steps
    mesh_desc = r.ctx.mesh_descriptor(mesh)       # the same host arrays every step, like the drop-in's TriangleMesh vectors

    def e2e_step():
        """Host buffers in, host buffer out, through the C ABI."""
        t_a = time.perf_counter()
        r.ctx.upload_mesh_descriptor(0, mesh_desc)                  # H2D: what UpdateTransforms produced this frame
        t_b = time.perf_counter()
        e2e_parts[0] += t_b - t_a
        if world == 1:
            tm = r.render_host_ptr(host_frame.data_ptr(), WIDTH * 4)     # kernel + progressive D2H, blocking
            e2e_parts[1] += time.perf_counter() - t_b
            e2e_parts[3] += tm["kernel_ms"] * 1e-3; e2e_parts[4] += tm["total_ms"] * 1e-3; e2e_parts[5] += 1
            return
        if present_mode == "direct":
            tm = r.render_strips_to_host(rank, world, surface.ptr, surface.pitch_bytes)   # blocking: this rank's strips are in host memory
            t_c = time.perf_counter()
            surface.arrive_and_wait(lib=r.ctx.lib)  # surface complete; nobody starts the next frame earlier
            e2e_parts[1] += t_c - t_b; e2e_parts[2] += time.perf_counter() - t_c
            e2e_parts[3] += tm["kernel_ms"] * 1e-3; e2e_parts[4] += tm["total_ms"] * 1e-3; e2e_parts[5] += 1
            return
        if signals:
            # every rank's CTAs bump per-band counters in rank 0's frame; rank 0 presents band after band
            present_no[0] += 1
            r.render_strips_to_frame_banded(rank, world, frame_ptr, PRESENT_BANDS, stream)
            if rank == 0:
                r.frame_present(host_frame.data_ptr(), WIDTH * 4, PRESENT_BANDS, present_no[0])
            dist.all_reduce(token)            # nobody starts the next frame before the root has presented this one
            torch.cuda.synchronize()
            return
        device_step()
        torch.cuda.synchronize()
        if rank == 0:
            if fused:
                r.download_to(host_frame.data_ptr(), WIDTH * 4)
            else:
                host_frame.copy_(frame_dev)

    # ---- device-timed region ---------------------------------------------------------------------
    def timed_device_region(sampler):
        poison_device_frame()
        for _ in range(max(args.warmup, 3)):
            device_step()
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        kstarts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        kends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        with sampler:
            for i in range(args.steps):
                flush.zero_()                      # evict the frame buffer from L2 between timed steps
                starts[i].record()
                kstarts[i].record()
                render_step()
                kends[i].record()
                exchange_step()
                ends[i].record()
            barrier()
        step_ms = torch.tensor([a.elapsed_time(b) for a, b in zip(starts, ends)], dtype=torch.float64, device="cuda")
        kern_ms = torch.tensor([a.elapsed_time(b) for a, b in zip(kstarts, kends)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)     # per step, the slowest rank
            dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
        steps_sorted = sorted(float(x) for x in step_ms)
        stats = {"min": steps_sorted[0], "median": steps_sorted[len(steps_sorted) // 2], "max": steps_sorted[-1]}
        return float(step_ms.sum()), float(kern_ms.mean()), stats, device_frame_check()

    sampler = ClockSampler(local_rank)
    r.ctx.set_mesh_path(PATHS["slab_linear"])
    slab_total_ms, slab_kernel_ms, slab_stats, slab_check = timed_device_region(ClockSampler(local_rank))
    r.ctx.set_mesh_path(0)                                     # default: BVH, the fixture carries the reference's nodes
    total_ms, kernel_ms, step_stats, device_check = timed_device_region(sampler)

    # ---- end-to-end region -----------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        if rank == 0:                             # the host surface starts poisoned, like the device frames
            (shared_surface if shared_surface is not None else host_frame.numpy().view(np.uint32))[...] = POISON
        barrier()
        for _ in range(max(args.warmup, 3)):
            e2e_step()
        e2e_parts[:] = [0.0] * 6                  # the first call pins the surface: not part of the per-step picture
        barrier()
        sampler2 = ClockSampler(local_rank)
        per_step = []
        with sampler2:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ta = time.perf_counter()
                e2e_step()
                per_step.append((time.perf_counter() - ta) * 1e3)
            barrier()
            t1 = time.perf_counter()
        sampler.merge(sampler2)
        e2e_s = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        per_step_t = torch.tensor(per_step, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
            dist.all_reduce(per_step_t, op=dist.ReduceOp.MAX)
        e2e_ms = float(e2e_s) * 1e3 / args.steps
        per_step = sorted(float(x) for x in per_step_t)
        h2d = int(mesh.positions.nbytes + mesh.indices.nbytes + mesh.normals.nbytes) * world
        e2e = {"value": RAYS_PER_FRAME / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": WIDTH * HEIGHT * 4,
               "step_ms": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]},
               "timing": "host wall clock around the blocking C-ABI calls of every step (upload + render into host memory), max over ranks"}
        if e2e_parts[5] > 0:
            # where a step's time goes, slowest rank per component, averaged over the timed steps
            parts = torch.tensor([x / e2e_parts[5] * 1e3 for x in e2e_parts[:5]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(parts, op=dist.ReduceOp.MAX)
            e2e["breakdown_ms_max_over_ranks"] = dict(zip(["host_upload_mesh", "host_render_strips_to_host_call", "host_wait_for_all_ranks",
                                                           "device_kernel", "device_kernel_and_copies"], [float(x) for x in parts]))

    # ---- in-process leg (N > 1): ONE context over all N devices on rank 0's process - the path under the reference's
    # `pRenderer->Render(pScene)` (source/main.cpp:91) when the drop-in is given RT_B200_DEVICES.  The other ranks wait
    # in a host-side (gloo) barrier with idle GPUs.
    in_process = None
    if world > 1 and not args.no_e2e:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                rr = Renderer(WIDTH, HEIGHT, device_ids=list(range(world)))
                rr.SetScene(scene)
                surface_ip = torch.empty((HEIGHT, WIDTH), dtype=torch.int32).pin_memory()
                surface_ip.numpy().view(np.uint32)[...] = POISON
                for _ in range(max(args.warmup, 3)):
                    rr.ctx.upload_mesh(0, mesh)
                    rr.render_host_ptr(surface_ip.data_ptr(), WIDTH * 4)
                ms_ip, k_ip = [], []
                for _ in range(args.steps):
                    ta = time.perf_counter()
                    rr.ctx.upload_mesh(0, mesh)
                    tm = rr.render_host_ptr(surface_ip.data_ptr(), WIDTH * 4)
                    ms_ip.append((time.perf_counter() - ta) * 1e3)
                    k_ip.append(tm["kernel_ms"])
                in_process = {"what": f"rt_create over devices 0..{world - 1} in one process, rt_upload_mesh + rt_render into pinned host memory per step (direct present: every device copies its own strips)",
                              "ms_per_step": statistics.mean(ms_ip), "step_ms": {"min": min(ms_ip), "median": statistics.median(ms_ip), "max": max(ms_ip)},
                              "Mrays_per_s": RAYS_PER_FRAME / (statistics.mean(ms_ip) * 1e-3) / 1e6, "kernel_ms_max_over_devices": statistics.mean(k_ip),
                              "devices": rr.ctx.device_count, "frame_check": frame_check_of(surface_ip.numpy().view(np.uint32))}
                rr.close()
            except Exception as exc:                   # noqa: BLE001  (reported; the headline legs are already measured)
                in_process = {"error": str(exc)}
        dist.barrier(group=host_group)

    if rank != 0:
        if world > 1:
            dist.barrier()
            if fused:
                r.frame_release(frame_ptr)
            dist.destroy_process_group()
        return 0

    # ---- the frame the end-to-end leg produced, against the frame the reference rendered ------------
    produced = (shared_surface if shared_surface is not None else host_frame.numpy().view(np.uint32)) if e2e is not None else None
    frame_check = frame_check_of(produced) if produced is not None else None

    # ---- roofline (rank 0's GPU) -----------------------------------------------------------------
    peak_nofma = r.ctx.measure_fp32_peak(False)
    peak_fma = r.ctx.measure_fp32_peak(True)
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    def roofline_of(pname, k_ms, kernel_name):
        # strips are dealt round-robin: every launch does ~1/N of the frame's work
        flop_per_launch = flop_per_frame[pname] / world
        achieved = flop_per_launch / (k_ms * 1e-3) / 1e12
        return {
            "bound": "fp32", "achieved": achieved, "peak": peak_nofma["tflops"], "unit": "TFLOP/s",
            "frac": achieved / peak_nofma["tflops"], "traffic": traffic if pname == "bvh" else None,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel, read from profiles/roofline_traffic.json (NOT measured in this run)" if pname == "bvh" else None,
            "frac_definition": "EXECUTED work: SURVEY.md 8(d) FLOP weights x the events THIS mesh body executes (counters build of the same kernel) / kernel time / peak",
            "peak_source": "measured live on this GPU (rt_measure_fp32_peak): FMUL+FADD issue peak; the reference's arithmetic is unfused, so FFMA is not available to this path",
            "peak_fma": peak_fma["tflops"], "frac_of_fma_peak": achieved / peak_fma["tflops"],
            "algorithmic_gflop_per_frame": flop_per_frame[pname] / 1e9,
            "flop_definition": "SURVEY.md 8(d) weights x the event counts of this mesh path (counters build of the same kernel)",
            "kernel": kernel_name, "kernel_ms": k_ms,
            "hbm": {"bytes_per_launch": WIDTH * HEIGHT * 4 // world, "achieved_gbs": WIDTH * HEIGHT * 4 / world / (k_ms * 1e-3) / 1e9,
                    "peak_gbs": _measured_peaks().get("hbm_gbs")},
        }

    roofline = roofline_of("bvh", kernel_ms, "rt::render_kernel_persistent<Combined, shadows, BVH>")
    # the same kernel time against the north-star algorithm's FLOP count (what a brute-force kernel would have to do)
    roofline["survey_algorithmic_gflop_per_frame"] = flop_per_frame["slab_linear"] / 1e9
    roofline["survey_algorithmic_tflops_equivalent"] = flop_per_frame["slab_linear"] / world / (kernel_ms * 1e-3) / 1e12
    # SURVEY.md 8(d) read literally: the brute-force (slab + every triangle) FLOP count over THIS kernel's time.  The BVH
    # body skips most of that work, so this can exceed 1; `frac` above counts only what the kernel executes, and
    # north_star_slab_linear.roofline.frac is the kernel that really runs 8(d)'s algorithm.
    roofline["frac_survey_8d"] = roofline["survey_algorithmic_tflops_equivalent"] / peak_nofma["tflops"]
    slab_ms_per_step = slab_total_ms / args.steps
    north_star = {"value": RAYS_PER_FRAME / (slab_ms_per_step * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": slab_ms_per_step,
                  "step_ms": slab_stats, "frame_check_device_leg": slab_check,
                  "roofline": roofline_of("slab_linear", slab_kernel_ms, "rt::render_kernel_persistent<Combined, shadows, slab + linear>")}

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        res, frames, kind = bounded_reference_run(100, 1, budget_s=15.0)
        ms = statistics.mean(res["ms"])
        one = run_reference(1, 0, threads=1) if kind == "reference" else None     # also the 1-thread time
        cpu_baseline = {"value": RAYS_PER_FRAME / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": res["threads"],
                        "kind": kind, "ms_per_frame": ms, "ms_per_frame_1_thread": one["ms_median"] if one else None,
                        "sample": f"{frames} full {WIDTH}x{HEIGHT} frames of the reference's Renderer::Render ({res.get('path')})"}

    # ---- the other BASELINE.json configs (640x480), N = 1 only: parity + times, not the headline ---------
    other = None
    if world == 1:
        other = {}
        for label, fixture, mode_shadows in (("Scene_W1 640x480 no shadows", "w1_640", (3, False)), ("Scene_W3 640x480", "w3_640", (3, True)),
                                             ("Scene_W4_ReferenceScene 640x480", "w4ref_640", (3, True)), ("Scene_W4_BunnyScene 640x480", "bunny_640", (3, True)),
                                             ("Scene_W4_OptionalScene 320x240 (3082 triangles)", "optional_320", (3, True)),
                                             ("Scene_W4_OptionalScene 640x480 (3082 triangles)", "optional_640", (3, True)),
                                             ("Scene_W4_OptionalScene 3840x2160 (3082 triangles)", "optional_4k", (3, True)),
                                             ("Scene_W2 3840x2160", "w2_4k", (3, True)), ("Scene_W3 3840x2160", "w3_4k", (3, True)),
                                             ("Scene_W4_ReferenceScene 3840x2160", "w4ref_4k", (3, True))):
            sc = load_rtsc(os.path.join(ROOT, "tests", "golden", fixture + ".rtsc"))
            rr = Renderer(sc.width, sc.height, device_ids=[local_rank])
            if not mode_shadows[1]:
                rr.ToggleShadows()
            rr.SetScene(sc)
            with open(os.path.join(ROOT, "tests", "golden", fixture + ".frame.xz"), "rb") as f:
                pl = np.frombuffer(lzma.decompress(f.read()), dtype=np.uint8).reshape(3, sc.height, sc.width).astype(np.uint32)
            want = (pl[0] << 16) | (pl[1] << 8) | pl[2]
            host = torch.empty((sc.height, sc.width), dtype=torch.int32).pin_memory()
            for _ in range(5):
                rr.render_host_ptr(host.data_ptr(), sc.width * 4)
            k_ms = [rr.render_device()["kernel_ms"] for _ in range(30)]
            t0 = time.perf_counter()
            for _ in range(30):
                rr.render_host_ptr(host.data_ptr(), sc.width * 4)
            e_ms = (time.perf_counter() - t0) / 30 * 1e3
            n_rays = rays(rr.count_frame(mesh_path=1))
            other[label] = {"kernel_ms": float(np.mean(k_ms)), "e2e_ms": e_ms, "rays_per_frame": n_rays,
                            "Mrays_per_s_kernel": n_rays / (float(np.mean(k_ms)) * 1e-3) / 1e6,
                            "differing_pixels_vs_reference_frame": int((host.numpy().view(np.uint32) != want).sum())}
            if fixture in ("optional_640", "optional_4k"):
                # N3 at the BASELINE sizes: roofline of the executed work and the reference's own CPU loop beside it
                flop = algorithmic_flops(rr.count_frame(mesh_path=2), 3)
                ach = flop / (float(np.mean(k_ms)) * 1e-3) / 1e12
                other[label]["roofline"] = {"bound": "fp32", "achieved": ach, "peak": peak_nofma["tflops"], "unit": "TFLOP/s", "frac": ach / peak_nofma["tflops"],
                                            "algorithmic_gflop_per_frame": flop / 1e9, "frac_definition": "executed work of the BVH body, as for the headline"}
                if not args.no_cpu_baseline and reference_available():
                    n_frames = 3 if fixture == "optional_640" else 1
                    ref = run_reference_scene("W4_Optional", sc.width, sc.height, n_frames, 0)
                    ref_ms = statistics.mean(ref["ms"])
                    other[label]["cpu_baseline"] = {"value": n_rays / (ref_ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": ref["threads"], "kind": "reference",
                                                    "ms_per_frame": ref_ms, "sample": f"{n_frames} full {sc.width}x{sc.height} frame(s) of Renderer::Render, no warm-up"}
            rr.close()

        # N1: untransformed mesh uploaded once, 64 bytes of transform per frame, vertices / normals
        # transformed on the device, slab + linear body (no BVH)
        try:
            from gp1_raytracer_2223_b200.scene_file import load_rtms
            sc = load_rtsc(os.path.join(ROOT, "tests", "golden", "bunny_320_yaw10.rtsc"))
            src = load_rtms(os.path.join(ROOT, "tests", "golden", "bunny_320_yaw10.rtms"))[0]
            rr = Renderer(WIDTH, HEIGHT, device_ids=[local_rank])
            rr.SetScene(sc)
            rr.ctx.upload_mesh_source(0, src.positions, src.indices, src.normals, sc.meshes[0].cull_mode, sc.meshes[0].material_index)
            host = torch.empty((HEIGHT, WIDTH), dtype=torch.int32).pin_memory()
            for _ in range(5):
                rr.ctx.transform_mesh(0, src.transform)
                rr.render_host_ptr(host.data_ptr(), WIDTH * 4)
            t0 = time.perf_counter()
            for _ in range(30):
                rr.ctx.transform_mesh(0, src.transform)
                tm = rr.render_host_ptr(host.data_ptr(), WIDTH * 4)
            other["device-side UpdateTransforms, bunny 3840x2160 (pose yaw 1.0), slab + linear body"] = {
                "e2e_ms": (time.perf_counter() - t0) / 30 * 1e3, "kernel_ms": tm["kernel_ms"], "h2d_bytes_per_frame": 64}
            rr.close()
        except Exception as exc:                       # noqa: BLE001  (reported, not fatal for the headline)
            other["device-side UpdateTransforms"] = {"error": str(exc)}

        # N1, second half: UpdateTransforms WITH BuildBVH on the device (rt_set_mesh_device_bvh): a new pose every frame,
        # one build per frame (two launches), BVH body.  Self-check without the CPU oracle: the last frame must equal the
        # frame of the transform-only path (slab + linear body) for the same pose, bit for bit.
        try:
            from gp1_raytracer_2223_b200.scene_file import load_rtmp
            sc = load_rtsc(os.path.join(ROOT, "tests", "golden", "bunny_320_steps3.rtsc"))
            steps = load_rtmp(os.path.join(ROOT, "tests", "golden", "bunny_320_steps3.rtmp"))[0]
            rr = Renderer(WIDTH, HEIGHT, device_ids=[local_rank])
            rr.SetScene(sc)
            rr.ctx.upload_mesh_source(0, steps.positions, steps.indices, steps.normals, sc.meshes[0].cull_mode, sc.meshes[0].material_index)
            rr.ctx.set_mesh_device_bvh(0, True)
            host = torch.empty((HEIGHT, WIDTH), dtype=torch.int32).pin_memory()

            def pose(k):
                a = np.float32(0.05 * (k + 1))
                c, s_ = np.float32(np.cos(a)), np.float32(np.sin(a))
                rot = np.array([[c, 0, -s_, 0], [0, 1, 0, 0], [s_, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)
                return (rot @ steps.transforms[0]).astype(np.float32)
            n_frames, t_total = 35, 0.0
            for k in range(n_frames):
                m = pose(k)
                t0 = time.perf_counter()
                rr.ctx.transform_mesh(0, m)
                tm = rr.render_host_ptr(host.data_ptr(), WIDTH * 4)
                if k >= 5:
                    t_total += time.perf_counter() - t0
            built = host.numpy().view(np.uint32).copy()
            rr.close()
            rr = Renderer(WIDTH, HEIGHT, device_ids=[local_rank])
            rr.SetScene(sc)
            rr.ctx.upload_mesh_source(0, steps.positions, steps.indices, steps.normals, sc.meshes[0].cull_mode, sc.meshes[0].material_index)
            rr.ctx.transform_mesh(0, pose(n_frames - 1))
            rr.render_host_ptr(host.data_ptr(), WIDTH * 4)
            other["device-side UpdateTransforms + BuildBVH, bunny 3840x2160, new pose every frame, BVH body"] = {
                "e2e_ms": t_total / (n_frames - 5) * 1e3, "kernel_ms": tm["kernel_ms"], "h2d_bytes_per_frame": 64,
                "differing_pixels_vs_slab_body_after_35_builds": int((built != host.numpy().view(np.uint32)).sum())}
            rr.close()
        except Exception as exc:                       # noqa: BLE001
            other["device-side UpdateTransforms + BuildBVH"] = {"error": str(exc)}

        # ... and the reference's own animated loop on this box's host cores: Scene::Update (RotateY + UpdateTransforms +
        # BuildBVH on the main thread, source/Scene.cpp:431-437, source/DataTypes.h:210-236) + Renderer::Render per frame
        if not args.no_cpu_baseline and reference_available():
            try:
                ref = run_reference_scene("W4_Bunny", WIDTH, HEIGHT, 10, 1, extra=["--animate", "0.016"])
                upd, ren = statistics.mean(ref["update_ms"]), statistics.mean(ref["ms"])
                other["reference animated loop, bunny 3840x2160: Scene::Update + Render per frame"] = {
                    "update_ms": upd, "render_ms": ren, "frame_ms": upd + ren, "cores": ref["threads"], "kind": "reference",
                    "sample": "10 frames, timer advanced by 16 ms per frame (a new mesh yaw every frame), 1 warm-up"}
                key = "device-side UpdateTransforms + BuildBVH, bunny 3840x2160, new pose every frame, BVH body"
                if key in other and "e2e_ms" in other[key]:
                    other[key]["vs_reference_animated_loop"] = (upd + ren) / other[key]["e2e_ms"]
            except Exception as exc:                   # noqa: BLE001
                other["reference animated loop"] = {"error": str(exc)}

    ms_per_step = total_ms / args.steps
    exchange = ("single GPU" if world == 1 else ("peer stores into rank 0's frame over NVLink (fused gather), completion by peer signal word + stream wait" if signals else
                ("peer stores into rank 0's frame over NVLink (fused gather) + NCCL barrier" if fused else "NCCL gather to rank 0 + unstripe")))
    line = {
        "metric": "Mrays/s", "value": RAYS_PER_FRAME / (ms_per_step * 1e-3) / 1e6, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": DATA,
        "config": bench_config(),
        "step_ms": step_stats,
        "partition": f"{bands.STRIP_ROWS}-row strips round-robin over {world} rank(s); " + exchange,
        "value_definition": "sum over the timed steps of (max over ranks of that step's CUDA-event time: kernel launch .. frame complete in rank 0's HBM)" +
                            ("" if world == 1 else "; ranks are NOT re-synchronised between steps (non-root ranks only signal, rank 0 waits for all signals of step k), so this is the frame-throughput of the free-running pipeline; the frame latency with back-pressure on every rank is the e2e leg"),
        "clocks": sampler.summary(),
        "e2e": e2e,
        "e2e_present": {"single": "one GPU: progressive present (copy follows the kernel band by band)",
                        "direct": "every rank copies its own strips into the shared host surface over its own PCIe link (rt_render_strips_to_host), band by band behind its kernel; completion by one flag word per rank in the same mapping; no peer traffic and no collective on this leg",
                        "gather": "strips stored into GPU 0's frame over NVLink, GPU 0 presents band by band (one PCIe link)"}[present_mode],
        "frame_check": frame_check,
        "frame_check_device_leg": device_check,
        "in_process_multi_device": in_process,
        "gpu_launches": args.steps * (world + ((world - 1) if signals else (0 if fused or world == 1 else 1))),
        "library": library,
        "mesh_path": "bvh (reference's shipped IntersectionTest_BVH over the reference's own nodes)",
        "kernel_variant": KERNEL_NOTE,
        "roofline": roofline,
        "north_star_slab_linear": north_star,
        "other_configs": other,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


if __name__ == "__main__":
    sys.exit(main())
