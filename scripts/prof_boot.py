"""Profiling driver: one bootstrap of one ciphertext at the reference ring between cudaProfilerStart/Stop."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import CKKS
c = CKKS(logN=15, L=28, dnum=4, sparse_h=192)
c.keygen(); c.gen_mult_key()
n = c.N // 2
c.bootstrap_setup((3, 3), n); c.bootstrap_keygen(n)
ct = c.encrypt(np.random.default_rng(0).uniform(-1, 1, n), level=24)
for _ in range(2): r = c.bootstrap(ct)
c.sync()
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
r = c.bootstrap(ct); c.sync()
rt.cudaProfilerStop()
print("done")
