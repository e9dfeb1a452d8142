import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from fhe_linformer_b200 import Engine
e = Engine(device=0, logN=16)
l, G = 28, 8
rng = np.random.default_rng(0)
N = e.N
evk = e.to_dev(rng.integers(0, 1 << 50, (e.dnum, 2, e.L + e.K, N), dtype=np.uint64))
g = e.galois(1)
pins = [torch.empty((G, 2, l, N), dtype=torch.int64).pin_memory() for _ in range(4)]
h = [p.numpy().view(np.uint64) for p in pins]
h[0][...] = rng.integers(0, 1 << 50, h[0].shape, dtype=np.uint64); h[1][...] = h[0]
for wait in (True, False):
    e.host_rotate_batch(h[0], g, evk, out=h[2], wait=True)
    t0 = time.perf_counter(); n = 10
    for i in range(n):
        e.host_rotate_batch(h[i % 2], g, evk, out=h[2 + i % 2], wait=wait)
    e.sync(); dt = time.perf_counter() - t0
    print("chunk %s wait=%s: %.0f rot/s  (%.1f GB/s each way)" % (os.environ.get("FLK_HOST_CHUNK", "2"), wait, n * G / dt, n * G * 2 * l * N * 8 / dt / 1e9))
# plain copies for reference
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d = torch.empty((G, 2, l, N), dtype=torch.int64, device="cuda"); d2 = torch.empty_like(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10):
    with torch.cuda.stream(s1): d.copy_(pins[0], non_blocking=True)
torch.cuda.synchronize(); print("H2D alone %.1f GB/s" % (10 * d.numel() * 8 / (time.perf_counter() - t0) / 1e9))
t0 = time.perf_counter()
for i in range(10):
    with torch.cuda.stream(s1): d.copy_(pins[0], non_blocking=True)
    with torch.cuda.stream(s2): pins[2].copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("H2D + D2H together %.1f GB/s each" % (10 * d.numel() * 8 / (time.perf_counter() - t0) / 1e9))
