"""Profiling driver: one 128-diagonal BSGS product (packed-layer pattern) at a given level between cudaProfilerStart/Stop."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import CKKS
level = int(sys.argv[1]) if len(sys.argv) > 1 else 0
c = CKKS(logN=15, L=28, dnum=4, sparse_h=192)
c.keygen(); c.gen_mult_key()
n = c.N // 2
rng = np.random.default_rng(0)
lt = c.linear_transform({128 * k: rng.uniform(-1, 1, n) for k in range(128)}, n)
c.gen_rot_keys(lt.rotations())
ct = c.encrypt(rng.uniform(-1, 1, n), level=level)
for _ in range(2): r = lt.apply(ct)
c.sync()
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
r = lt.apply(ct); c.sync()
rt.cudaProfilerStop()
print("done")
