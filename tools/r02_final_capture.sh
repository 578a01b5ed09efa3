#!/bin/bash
# Round-2 evidence on one B200 (through gpurun): bench line, reference arm, ncu launch list of the bench, per-launch times of
# the wavefront path, ncu --set full of its two walk kernels.
O=gpurun_out/r02_final; mkdir -p $O
python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 200 $O/bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference_arm.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
RT_B200_WAVE_TIMING=1 python tools/prof_target.py optional_320 2> $O/wave_timing_optional_320.txt > /dev/null
RT_B200_WAVE_TIMING=1 python tools/prof_target.py optional_640 2> $O/wave_timing_optional_640.txt > /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"view_walk|shadow_walk" --launch-skip 4 -c 2 -o $O/prof_wave_walks -f python tools/prof_target.py optional_320 > /dev/null 2>&1
python tools/wave_crossover.py optional_320 > $O/wave_crossover.txt 2>&1
python tools/share_e2e.py 1 2 4 8 > $O/share_e2e.txt 2>&1
python tools/band_cost.py > $O/band_cost.txt 2>&1
python tools/copy_under_kernel.py > $O/copy_under_kernel.txt 2>&1
ls -la $O
