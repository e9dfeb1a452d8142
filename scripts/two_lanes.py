"""Do two independent batched EvalRotate streams overlap usefully on one GPU?  (DRAM-bound passes of one lane under the
multiplier-bound passes of the other.)  Two Engine objects = two streams + two pools; same total work as one lane."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhe_linformer_b200 import Engine
logN, l = 16, 28
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lanes = [Engine(device=0, logN=logN) for _ in range(2)]
rng = np.random.default_rng(0)
N = lanes[0].N
def operands(e):
    ct = np.stack([np.stack([rng.integers(0, int(e.moduli[m]), N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
    cts = e.to_dev(np.broadcast_to(ct, (B,) + ct.shape).copy()); out = e.buf(cts.shape)
    evk = e.to_dev(rng.integers(0, 1 << 50, (e.dnum, 2, e.L + e.K, N), dtype=np.uint64))
    return cts, out, evk
ops = [operands(e) for e in lanes]
g = lanes[0].galois(1)
def run(which, reps):
    for e in lanes: e.sync()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        for k in which:
            cts, out, evk = ops[k]
            lanes[k].rotate_batch(cts, g, evk, out=out)
    for e in lanes: e.sync()
    return (time.perf_counter() - t) / (reps * len(which) * B)
for _ in range(2): run([0, 1], 2)
one = run([0, 0], 10)
two = run([0, 1], 10)
print("B=%d per call: one lane %.1f us/rotation (%.0f rot/s), two lanes alternating %.1f us/rotation (%.0f rot/s)" % (B, one * 1e6, 1 / one, two * 1e6, 1 / two))
