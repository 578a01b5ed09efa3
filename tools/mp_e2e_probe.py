#!/usr/bin/env python
"""Phase timing of the multi-process banded present (torchrun --nproc-per-node N tools/mp_e2e_probe.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from gp1_raytracer_2223_b200 import Renderer, load_rtsc
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scene = load_rtsc(os.path.join(ROOT, "tests", "golden", "bunny_4k.rtsc"))
r = Renderer(3840, 2160, device_ids=[lr]); r.SetScene(scene)
h = [r.frame_export() if rank == 0 else None]; dist.broadcast_object_list(h, src=0)
ptr = 0 if rank == 0 else r.frame_import(h[0])
host = torch.empty((2160, 3840), dtype=torch.int32).pin_memory() if rank == 0 else None
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
token = torch.zeros(1, dtype=torch.int32, device="cuda")
B = int(os.environ.get("BANDS", "16"))
n = 0
acc = np.zeros(4)
for it in range(60):
    n += 1
    t0 = time.perf_counter()
    MODE = os.environ.get("MODE", "banded")
    if MODE == "banded": r.render_strips_to_frame_banded(rank, world, ptr, B, stream)
    else: r.render_strips_to_frame(rank, world, ptr, stream)
    t1 = time.perf_counter()
    if rank == 0 and MODE == "banded": r.frame_present(host.data_ptr(), 3840 * 4, B, n)
    t2 = time.perf_counter()
    dist.all_reduce(token); torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it >= 10: acc += [t1 - t0, t2 - t1, t3 - t2, t3 - t0]
print(f"rank {rank}: launch {acc[0]/50*1e3:.3f} ms, present {acc[1]/50*1e3:.3f} ms, barrier+sync {acc[2]/50*1e3:.3f} ms, total {acc[3]/50*1e3:.3f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
