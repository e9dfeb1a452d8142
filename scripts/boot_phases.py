import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from fhe_linformer_b200 import CKKS
c = CKKS(logN=15, L=28, dnum=4)
n = 16384
c.keygen(5); c.gen_mult_key(); c.bootstrap_setup((3,3), n); c.bootstrap_keygen(n); c.sync()
rng = np.random.default_rng(0)
vs = [rng.uniform(-1,1,n) for _ in range(5)]
cts = [c.encrypt(v, level=24) for v in vs]
c.lib.fl_batch_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
arr = (C.c_void_p*5)(*[e.h for e in cts]); b5 = c._out(c.lib.fl_batch_pack, arr, 5)
for name, x in (("single", cts[0]), ("batch5", b5)):
    c.bootstrap(x); c.sync()
    print(name, file=sys.stderr, flush=True)
    c.bootstrap(x); c.sync()
