"""Algorithmic FLOP accounting of one frame.

The numerator of the FP32 roofline is defined on the north-star algorithm (reference order,
reference early exits, slab + every triangle), 1 FLOP per FP32 add/sub/mul/div/sqrt/min/max,
powf = 1, compares / selects / integer work = 0.  The event counts come from the counters
build of the kernel (rt_count_frame) - or, in tests, from the CPU oracle's identical counters.
"""
from __future__ import annotations

import numpy as np

# slot -> FLOP per event (slots: enum rt_counter_slot in include/rt_b200.h)
WEIGHTS = {
    0: 38 + 11,      # ray-gen incl. normalise and 1/dir; shadowFactor, MaxToOne, x255
    1: 6,            # offset origin per hit pixel
    2: 16, 3: 19, 4: 28, 5: 9,       # primary sphere tests
    6: 16, 7: 19, 8: 19,             # shadow sphere tests
    9: 14, 10: 6, 11: 14,            # plane tests
    12: 22, 14: 22,                  # slab tests
    16: 5, 17: 25, 18: 35, 19: 51, 20: 57, 21: 63,   # primary triangle exits
    22: 5, 23: 25, 24: 35, 25: 51, 26: 57, 27: 63,   # shadow triangle exits
    28: 15,          # light iteration setup
    32: 0, 33: 6, 34: 33, 35: 112,   # Shade per material class
    36: 22, 37: 22,                  # BVH path: node box tests (same arithmetic as the mesh slab test)
}
# per un-shadowed light evaluation (slot 31), by lighting mode
LIT_WEIGHT = {3: 27, 0: 9, 1: 15, 2: 3}


def algorithmic_flops(counters, lighting_mode: int = 3) -> int:
    c = np.asarray(counters, dtype=np.uint64)
    total = sum(int(c[k]) * w for k, w in WEIGHTS.items())
    total += int(c[31]) * LIT_WEIGHT[int(lighting_mode)]
    if lighting_mode in (0, 1):       # Shade is not evaluated in ObservedArea / Radiance modes
        total -= sum(int(c[k]) * WEIGHTS[k] for k in (33, 34, 35))
    return total


def rays(counters) -> int:
    """Primary rays + shadow rays cast (a shadow ray per light per hit pixel when shadows are on)."""
    c = np.asarray(counters, dtype=np.uint64)
    return int(c[0]) + int(c[29])
