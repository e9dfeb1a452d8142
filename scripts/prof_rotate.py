"""Profiling driver: `reps` batched EvalRotate calls at full chain (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import Engine
logN = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
l = int(sys.argv[3]) if len(sys.argv) > 3 else 28
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
e = Engine(device=0, logN=logN)
rng = np.random.default_rng(0)
N = e.N
ct = np.stack([np.stack([rng.integers(0, int(e.moduli[m]), N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
cts = e.to_dev(np.stack([np.roll(ct, i, axis=2) for i in range(B)]))
evk = e.to_dev(rng.integers(0, 1 << 50, (e.dnum, 2, e.L + e.K, N), dtype=np.uint64))
out = e.buf(cts.shape)
for _ in range(reps):
    e.rotate_batch(cts, e.galois(1), evk, out=out)
e.sync()
print("done")
