"""Development probe: the stand-alone 63-bit pair sort at N keys (for ncu).  usage: sort_probe.py N"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parallelnbody_b200 as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
keys = np.random.default_rng(1).integers(0, 1 << 63, n, dtype=np.uint64)
_, _, ms = P.sort_pairs_u64(keys, 63, timed=True)
print(f"N={n}: {ms:.3f} ms = {n / ms * 1e-6:.2f} Gkeys/s, {ms / 8 * 1e3:.1f} us per pass = {n * 24 / (ms / 8) * 1e-6:.0f} GB/s")
