import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from fhe_linformer_b200 import CKKS
c = CKKS(logN=13, L=24, dnum=4, sparse_h=64)
c.keygen(); c.gen_mult_key()
n = c.N // 2
c.bootstrap_setup((3, 3), n); c.bootstrap_keygen(n)
rng = np.random.default_rng(1)
for amp in (1.0, 0.01):
    v = rng.uniform(-amp, amp, n)
    for lvl in (c.L - 3, c.L - 2, c.L - 1):
        ct = c.encrypt(v, level=lvl)
        b = c.bootstrap(ct)
        print("amp %.2f deg1 level %d (l=%d): out level %d err %.2e" % (amp, lvl, c.L - lvl, b.level, np.abs(c.decrypt(b) - v).max()))
    for lvl in (c.L - 3, c.L - 2):
        ct = c.mult(c.encrypt(v, level=lvl), c.encode(np.ones(n), level=lvl))
        try:
            b = c.bootstrap(ct)
            print("amp %.2f deg2 level %d (l=%d): out level %d err %.2e" % (amp, lvl, c.L - lvl, b.level, np.abs(c.decrypt(b) - v).max()))
        except Exception as e:
            print("amp %.2f deg2 level %d: %s" % (amp, lvl, e))
