#!/bin/bash
# wavefront: when to take subtrees in parts (RT_B200_WAVE_PARTS_BELOW: jobs per SM below which a walk kernel does) and
# in how many (RT_B200_WAVE_PARTS), kernel ms on the optional scene
for parts in 4 2; do
for pb in 0 60 110 200 100000; do
  echo "== RT_B200_WAVE_PARTS=$parts RT_B200_WAVE_PARTS_BELOW=$pb"
  RT_B200_WAVE_PARTS=$parts RT_B200_WAVE_PARTS_BELOW=$pb python tools/wave_ab.py optional_320 optional_640 2>&1 | grep "wavefront"
done
done
RT_B200_WAVE_TIMING=1 python tools/wave_ab.py optional_320 optional_640 2>&1 | grep "^wave: primary" | sort | uniq -c | sort -rn | head -6
