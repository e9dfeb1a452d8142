import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import CKKS
logN = int(sys.argv[1]) if len(sys.argv) > 1 else 12
L = int(sys.argv[2]) if len(sys.argv) > 2 else 19
h = int(sys.argv[3]) if len(sys.argv) > 3 else 192
c = CKKS(logN=logN, L=L, dnum=4 if L >= 20 else 3, sparse_h=h)
n = c.N // 2
t0 = time.time(); c.keygen(5); c.gen_mult_key(); c.bootstrap_setup((3, 3), n); t1 = time.time(); c.bootstrap_keygen(n); c.sync(); t2 = time.time()
print("setup %.2fs keygen %.2fs rot keys %d" % (t1 - t0, t2 - t1, c.num_rot_keys()))
rng = np.random.default_rng(0)
for start_level, amp in [(L - 3, 1.0), (L - 2, 0.25), (L - 4, 1.0)]:
    v = rng.uniform(-amp, amp, n)
    ct = c.encrypt(v, level=start_level)
    t0 = time.time(); b = c.bootstrap(ct); c.sync(); dt = time.time() - t0
    out = c.decrypt(b)
    print("in level %d -> out level %d deg %d limbs %d | max err %.3e | %.1f ms" % (ct.level, b.level, b.deg, b.limbs, np.abs(out - v).max(), dt * 1e3))
# bootstrap then keep computing
m = c.mult(b, b); print("post-boot mult err %.3e level %d" % (np.abs(c.decrypt(m) - v * v).max(), m.level))
