"""NTT / INTT throughput alone (algorithmic GB/s = 16 N per limb), full-chain limbs, working set > L2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhe_linformer_b200 import Engine
logN = int(sys.argv[1]) if len(sys.argv) > 1 else 16
e = Engine(device=0, logN=logN)
N, l, B = e.N, 28, 16
rng = np.random.default_rng(0)
base = np.stack([rng.integers(0, int(e.moduli[m % l]), N, dtype=np.uint64) for m in range(2 * l)])
midx = np.concatenate([np.arange(l), np.arange(l)]).astype(np.int32)
bufs = [e.to_dev(np.roll(base, i, axis=1)) for i in range(B)]
stream = torch.cuda.ExternalStream(e.stream())
for name, fn in (("ntt", e.ntt), ("intt", e.intt)):
    for b in bufs: fn(b, midx)
    e.sync()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(5):
        for b in bufs: fn(b, midx)
    z.record(stream); e.sync()
    ms = a.elapsed_time(z) / (5 * B)
    print("%-5s N=2^%d  %d limbs: %.1f us  %.0f GB/s algorithmic (%.1f%% of 6549)" % (name, logN, 2 * l, ms * 1e3, 16 * N * 2 * l / ms / 1e6, 16 * N * 2 * l / ms / 1e6 / 65.49))
