"""Profiling driver: one batched rotsum ladder (7 doubling steps, stride 128: the ladder of matmulRE, F.cpp:869-883) at the reference
parameters, for ncu launch lists.  usage: prof_rotsum.py [level] [batch] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import CKKS
level = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
c = CKKS(logN=15, L=28, dnum=4)
c.keygen(3)
import ctypes as C
rots = (C.c_int * 64)()
nr = c.lib.fl_rotsum_rotations(7, 128, rots, 64)          # doubling keys + the extra multiples of the hoisted groups
c.gen_rot_keys([int(rots[i]) for i in range(nr)])
n = c.N // 2
rng = np.random.default_rng(0)
vs = [rng.uniform(-1, 1, n) for _ in range(B)]
b = c.pack([c.encrypt(v, level=level) for v in vs])
rotsum = lambda x: c._out(c.lib.fl_rotsum, x.h, 7, 128)
c.lib.fl_rotsum.restype = C.c_int
c.lib.fl_rotsum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
r = rotsum(b); c.sync()
rt = C.CDLL("libcudart.so")
rt.cudaProfilerStart()          # under ncu --profile-from-start off: exactly one ladder
r = rotsum(b); c.sync()
rt.cudaProfilerStop()
t0 = time.time()
for _ in range(reps):
    r = rotsum(b)
c.sync()
dt = (time.time() - t0) / reps
out = c.decrypt(c.unpack(r)[0])
ref = vs[0].reshape(128, 128).sum(0)
print("rotsum(7, 128) level %d batch %d: %.2f ms per call (%.1f us per ciphertext-rotation), err %.2e" % (level, B, dt * 1e3, dt * 1e6 / (7 * B), np.abs(out[:128] - ref).max()))
