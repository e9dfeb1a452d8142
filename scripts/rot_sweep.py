"""Batched EvalRotate throughput over (logN, limbs, batch): rotations/s and fraction of the HBM roofline (developer script)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhe_linformer_b200 import Engine
for logN in (15, 16):
    e = Engine(device=0, logN=logN)
    N = e.N; rng = np.random.default_rng(0)
    evk = e.to_dev(rng.integers(0, 1 << 50, (e.dnum, 2, e.L + e.K, N), dtype=np.uint64))
    stream = torch.cuda.ExternalStream(e.stream())
    for l in (28, 20, 12, 5):
        for B in (1, 8, 64):
            if logN == 16 and B == 64 and l > 20: continue
            ct = np.stack([np.stack([rng.integers(0, int(e.moduli[m]), N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
            cts = e.to_dev(np.broadcast_to(ct, (B,) + ct.shape).copy()); out = e.buf(cts.shape)
            g = e.galois(1)
            for _ in range(2): e.rotate_batch(cts, g, evk, out=out)
            e.sync()
            reps = max(2, 256 // (B * max(1, l // 6)))
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps): e.rotate_batch(cts, g, evk, out=out)
            z.record(stream); e.sync()
            us = a.elapsed_time(z) * 1e3 / (reps * B)
            beta = -(-l // e.alpha)
            alg = (4 * l + 2 * beta * (l + e.K)) * 8 * N
            print("N=2^%d l=%2d B=%2d: %7.1f us/rotation  %7.0f rot/s  %5.1f%% of HBM roofline" % (logN, l, B, us, 1e6 / us, alg / us / 1e3 / 6549.1 * 100), flush=True)
            cts.free(); out.free()
    e.close()
