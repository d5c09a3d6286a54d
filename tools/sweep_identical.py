"""Upper bound for a work-queue kernel: identical instances (no iteration variance) at different block sizes (dev tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200
from tools.sweep import timeit
c = dart_b200.workloads.pmpc_config2(64)
i = 5
B = 131072
same = {k: np.repeat(np.asarray(v)[i:i + 1], B, axis=0) for k, v in c.items()}
for lanes in (4, 8):
    for bt in (32, 64, 128, 256):
        ms, ok, lc = timeit(dart_b200.pmpc_cfg(lanes=lanes, block_threads=bt), same, 3)
        print(f"identical B={B} lanes={lanes} bt={bt} ms={ms:.3f} solves/s={B/ms*1e3:.3e} grid={lc['grid']} smem={lc['smem_bytes']}", flush=True)
