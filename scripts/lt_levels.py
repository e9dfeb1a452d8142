"""Time of one 128-diagonal BSGS product (the packed layers' pattern: shifts 128 k) as a function of the level."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import CKKS
c = CKKS(logN=15, L=28, dnum=4, sparse_h=192)
c.keygen(); c.gen_mult_key()
n = c.N // 2
rng = np.random.default_rng(0)
diags = {128 * k: rng.uniform(-1, 1, n) for k in range(128)}
lt = c.linear_transform(diags, n)
c.gen_rot_keys(lt.rotations())
v = rng.uniform(-1, 1, n)
for level in (0, 1, 8, 16, 22):
    ct = c.encrypt(v, level=level)
    for _ in range(2): r = lt.apply(ct)
    c.sync()
    ts = []
    for _ in range(5):
        t = time.perf_counter(); r = lt.apply(ct); c.sync(); ts.append(time.perf_counter() - t)
    print("level %2d (%2d limbs): %.2f ms per product" % (level, c.L - level, 1e3 * sorted(ts)[2]), flush=True)
