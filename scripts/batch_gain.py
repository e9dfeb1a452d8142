"""How much cheaper per ciphertext the forward's building blocks get when B samples share a call (config-5 throughput mode)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import CKKS
c = CKKS(logN=15, L=28, dnum=4, sparse_h=192)
c.keygen(); c.gen_mult_key()
n = c.N // 2
rng = np.random.default_rng(0)
lt = c.linear_transform({128 * k: rng.uniform(-1, 1, n) for k in range(128)}, n)
c.gen_rot_keys(lt.rotations() + [128 << i for i in range(7)] + [1 << i for i in range(7)])
c.bootstrap_setup((3, 3), n); c.bootstrap_keygen(n)
c.lib.fl_batch_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
def pack(elems):
    arr = (C.c_void_p * len(elems))(*[e.h for e in elems]); return c._out(c.lib.fl_batch_pack, arr, len(elems))
def timed(f, reps=3):
    f(); c.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); c.sync(); ts.append(time.perf_counter() - t)
    return 1e3 * sorted(ts)[len(ts) // 2]
coef = np.polynomial.chebyshev.chebinterpolate(lambda x: np.tanh(3 * x), 119); coef[0] *= 2
for B in (1, 4, 8):
    row = []
    for name, level, f in (("lt@1", 1, lambda x: lt.apply(x)), ("lt@16", 16, lambda x: lt.apply(x)), ("cheb119@17", 17, lambda x: c.eval_chebyshev(x, coef, -1, 1)),
                           ("cheb119@8", 8, lambda x: c.eval_chebyshev(x, coef, -1, 1)), ("rotsum7@2", 2, lambda x: c.rotsum(x, 7, 128)), ("boot@24", 24, lambda x: c.bootstrap(x))):
        x = pack([c.encrypt(rng.uniform(-1, 1, n), level=level) for _ in range(B)]) if B > 1 else c.encrypt(rng.uniform(-1, 1, n), level=level)
        row.append("%s %.2f" % (name, timed(lambda: f(x)) / B))
    print("B=%d  ms per ciphertext: %s" % (B, "  ".join(row)), flush=True)
