#!/bin/bash
# bench.py at N = 2, 4, 8 on one 8-GPU box, the way the driver launches it (run through `gpurun --gpus 8`).
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/bench_${n}gpu_final2.json 2> gpurun_out/bench_${n}gpu_final2.err
done
python - <<'PY'
import json
for n in (2, 4, 8):
    try:
        j = json.loads(open(f"gpurun_out/bench_{n}gpu_final2.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(j["ms_per_step"], 4), "value", round(j["value"]), "e2e ms", round(j["e2e"]["ms_per_step"], 4), j["e2e"].get("breakdown_ms_max_over_ranks"), j["frame_check"], j["e2e_present"][:30])
    except Exception as exc:
        print(n, "failed", exc)
        print(open(f"gpurun_out/bench_{n}gpu_final2.err").read()[-1200:])
PY
