#!/bin/bash
# bench.py's end-to-end leg with 32 and 16 present bands, the calling thread's phases beside it (RT_B200_HOST_TIMING)
for b in 32 16; do
  echo "== bands $b"
  RT_B200_PIPELINE_BANDS=$b RT_B200_HOST_TIMING=1 python bench.py --no-cpu-baseline 2> gpurun_out/bench_probe_$b.err | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print(j['e2e']['ms_per_step'], j['e2e']['step_ms'], j['e2e']['breakdown_ms_max_over_ranks'])"
  grep "rt_render host" gpurun_out/bench_probe_$b.err | sed -n '50p;80p;100p'
done
