#!/bin/bash
# Multi-GPU bench lines like the driver's SCALE run: bash tools/r02_scale_session.sh "8 4 2" (through gpurun --gpus 8)
mkdir -p gpurun_out/r02_scale
for n in ${1:-8 4 2}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps ${2:-20} --warmup 5 > gpurun_out/r02_scale/bench_${n}gpu.json 2> gpurun_out/r02_scale/bench_${n}gpu.err
  tail -c 300 gpurun_out/r02_scale/bench_${n}gpu.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_scale/bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("N=$n value", round(j["value"]), "ms", round(j["ms_per_step"],4), j["step_ms"], "| e2e ms", round(j["e2e"]["ms_per_step"],4), j["e2e"].get("breakdown_ms_max_over_ranks"))
    print("   checks", j["frame_check"]["differing_pixels_vs_reference_frame"], j["frame_check_device_leg"]["differing_pixels_vs_reference_frame"], (j["in_process_multi_device"] or {}).get("frame_check"), "in-process ms", (j["in_process_multi_device"] or {}).get("ms_per_step"))
except Exception as e:
    print("N=$n failed", e)
PY
done
