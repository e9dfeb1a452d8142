"""Profiling driver: one Chebyshev evaluation (degree argv[1], default 300) of one ciphertext at the reference ring, bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` lists exactly its launches; prints the wall time of warm calls too."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import CKKS
deg = int(sys.argv[1]) if len(sys.argv) > 1 else 300
level = int(sys.argv[2]) if len(sys.argv) > 2 else 16
c = CKKS(logN=15, L=28, dnum=4, sparse_h=192)
c.keygen(); c.gen_mult_key()
n = c.N // 2
v = np.random.default_rng(1).uniform(-1, 1, n)
ct = c.encrypt(v, level=level)
f = lambda x: np.tanh(50 * x)
coef = np.polynomial.chebyshev.chebinterpolate(f, deg); coef[0] *= 2
for _ in range(2): r = c.eval_chebyshev(ct, coef, -1, 1)
c.sync()
ts = []
for _ in range(5):
    t = time.perf_counter(); r = c.eval_chebyshev(ct, coef, -1, 1); c.sync(); ts.append(time.perf_counter() - t)
print("degree %d at level %d: %.2f ms per evaluation (warm, wall incl. sync), out level %d" % (deg, level, 1e3 * sorted(ts)[2], r.level))
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
r = c.eval_chebyshev(ct, coef, -1, 1); c.sync()
rt.cudaProfilerStop()
print("err vs tanh: %.2e" % np.abs(c.decrypt(r) - f(v)).max())
