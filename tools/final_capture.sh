#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): GPU tests of the newest rows, the default bench line, the
# reference arm, the ncu launch list of the bench and one full capture each of the pixel kernel and the build kernels.
timeout 200 python -m pytest tests/test_obj_file.py -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r01_final2.json 2> gpurun_out/bench_r01_final2.err; tail -c 300 gpurun_out/bench_r01_final2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel_persistent --launch-skip 2 -c 1 -o gpurun_out/prof_persistent_final -f python tools/prof_target.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:update_transforms_bvh|build_subtrees' --launch-skip 100 -c 2 -o gpurun_out/prof_build_v3 -f python tools/build_time.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_r01_final2.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["e2e"], j["roofline"]["frac"], j["cpu_baseline"])
for k, v in j["other_configs"].items():
    print(k, v)
PY
