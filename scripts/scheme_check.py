import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, math
from oracle.oracle import Oracle
from fhe_linformer_b200 import CKKS
P = dict(logN=int(sys.argv[1]) if len(sys.argv) > 1 else 12, L=int(sys.argv[2]) if len(sys.argv) > 2 else 14, dnum=3)
o = Oracle(**P); c = CKKS(sparse_h=64, **P)
seed = 42
c.keygen(seed); c.gen_mult_key(); c.gen_rot_keys([1, -1, 4]); c.gen_conj_key()
sk = o.gen_sk(seed, h=64); pk = o.gen_pk(seed + 1, sk)
assert (c.export_sk() == sk).all(), "sk"; assert (c.export_pk() == pk).all(), "pk"
assert (c.export_evk(0) == o.gen_relin_key(seed + 2, sk)).all(), "relin"
g = o.galois(1); assert (c.export_evk(g) == o.gen_galois_key(seed + 1000 + g, sk, g)).all(), "galois key"
print("keys bit-exact")
n = o.N // 2; rng = np.random.default_rng(0)
v = rng.uniform(-1, 1, n); w = rng.uniform(-1, 1, n)
pt = c.encode(v, level=0)
assert (pt.export()[0] == o.encode(v, o.sf[0], o.L)).all(), "encode"
pt3 = c.encode(v + 1j * w, level=3, slots=n)
assert (pt3.export()[0] == o.encode(v + 1j * w, o.sf[3], o.L - 3)).all(), "encode lvl3"
ct = c.encrypt(pt, seed=7)
assert (ct.export() == o.encrypt(7, o.encode(v, o.sf[0], o.L), pk)).all(), "encrypt"
print("encode/encrypt bit-exact")
assert np.abs(c.decrypt(ct) - v).max() < 1e-8
ct2 = c.encrypt(w)
s = c.add(ct, ct2); assert np.abs(c.decrypt(s) - (v + w)).max() < 1e-8
m = c.mult(ct, ct2); print("mult lvl/deg", m.level, m.deg, np.abs(c.decrypt(m) - v * w).max())
mp = c.mult(ct, c.encode(w, level=0)); print("ptmult", mp.level, mp.deg, np.abs(c.decrypt(mp) - v * w).max())
# FLEXIBLEAUTO pattern of main.cpp: weight at x.level, bias at x.level+1
x = mp
wt = c.encode(w, level=x.level); y = c.mult(x, wt); print("chain", y.level, y.deg, np.abs(c.decrypt(y) - v * w * w).max())
b = c.encode(v, level=y.level); z = c.add(y, b); print("bias add", z.level, z.deg, np.abs(c.decrypt(z) - (v * w * w + v)).max())
r = c.rotate(z, 1); print("rot", np.abs(c.decrypt(r) - np.roll(v * w * w + v, -1)).max())
r = c.rotate(z, -1); print("rot-1", np.abs(c.decrypt(r) - np.roll(v * w * w + v, 1)).max())
cc = c.conjugate(c.encrypt(v + 1j * w)); print("conj", np.abs(c.decrypt(cc, complex_out=True) - (v - 1j * w)).max())
# adds across levels
a1 = c.add(ct, z); print("add across levels", a1.level, a1.deg, np.abs(c.decrypt(a1) - (v + v * w * w + v)).max())
mm = c.mult(z, ct); print("mult across levels", mm.level, mm.deg, np.abs(c.decrypt(mm) - (v * w * w + v) * v).max())
am = c.add_many([ct, ct2, ct, ct2, ct]); print("addmany", np.abs(c.decrypt(am) - (3 * v + 2 * w)).max())
mc = c.mult(ct, 0.37); print("mult const", mc.level, mc.deg, np.abs(c.decrypt(mc) - 0.37 * v).max())
ac = c.add(m, 0.5); print("add const", np.abs(c.decrypt(ac) - (v * w + 0.5)).max())
# polynomial evaluation
pe = c.eval_poly(ct, [1, 1, 1 / 2., 1 / 6., 1 / 24., 1 / 120., 1 / 720.])
ref = sum(cf * v ** i for i, cf in enumerate([1, 1, 1 / 2., 1 / 6., 1 / 24., 1 / 120., 1 / 720.]))
print("evalpoly deg6: levels used", pe.level - ct.level, "deg", pe.deg, "err", np.abs(c.decrypt(pe) - ref).max())
mm8 = c.mult_many([pe] * 8); print("multmany8: level", mm8.level, "err", np.abs(c.decrypt(mm8) - ref ** 8).max() / np.abs(ref ** 8).max())
t0 = time.time()
gl = c.eval_chebyshev_function(lambda x: 0.5 * x * 8 * (1 + math.erf(x * 8 / 1.41421356237)), ct, -1, 1, 119)
gref = np.array([0.5 * x * 8 * (1 + math.erf(x * 8 / 1.41421356237)) for x in v])
print("gelu cheb119: levels", gl.level - ct.level, "deg", gl.deg, "err", np.abs(c.decrypt(gl) - gref).max(), "time", time.time() - t0)
if o.L >= 14:
    th = c.eval_chebyshev_function(lambda x: math.tanh(x * 50), ct, -1, 1, 300)
    print("tanh cheb300: levels", th.level - ct.level, "err", np.abs(c.decrypt(th) - np.tanh(50 * v)).max())
    u = c.encrypt(rng.uniform(20, 100, n)); uv = c.decrypt(u)
    iv = c.eval_chebyshev_function(lambda x: 1 / x, u, -1, 128, 119)
    print("inv cheb119 on [-1,128]: levels", iv.level - u.level, "err", np.abs(c.decrypt(iv) - 1 / uv).max())
# save / load
c.save(z, "/tmp/z.bin"); z2 = c.load("/tmp/z.bin"); assert (z2.export() == z.export()).all() and z2.level == z.level
c.save_keys("/tmp/keys.bin"); c.clear_keys(0); assert c.num_rot_keys() == 0; c.load_keys("/tmp/keys.bin"); assert c.num_rot_keys() == 4
print("rot after reload", np.abs(c.decrypt(c.rotate(z, 4)) - np.roll(v * w * w + v, -4)).max())
print("SCHEME OK")
