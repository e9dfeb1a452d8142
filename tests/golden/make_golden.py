"""Generates tests/golden/small_ring.npz from the CPU oracle (N=1024, 6 Q limbs, 3 digits).

The reference ships no golden vectors for this path and OpenFHE cannot be built here, so these vectors
pin the oracle against itself across refactors (and the CUDA path against the oracle on the GPU box,
where /root/reference and a compiler may be unavailable)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.oracle import Oracle  # noqa: E402

o = Oracle(logN=10, L=6, dnum=3)
rng = np.random.default_rng(20261018)
q = [int(x) for x in o.moduli]
poly = np.stack([rng.integers(0, q[m], o.N, dtype=np.uint64) for m in range(o.L)])
ct = np.stack([np.stack([rng.integers(0, q[m], o.N, dtype=np.uint64) for m in range(o.L)]) for _ in range(2)])
seed_sk, seed_evk, seed_rk = 11, 12, 13
sk = o.gen_sk(seed_sk, h=64)
g = o.galois(1)
rot = o.rotate(ct, g, o.gen_galois_key(seed_evk, sk, g))
mult = o.mul_relin(ct, ct, o.gen_relin_key(seed_rk, sk))
resc = np.stack([o.rescale(mult[0]), o.rescale(mult[1])])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_ring.npz"), moduli=o.moduli, roots=o.roots,
                    poly_coeff=poly, poly_eval=o.ntt(poly), ct=ct, rotated=rot, mult=mult, rescaled=resc,
                    seed_sk=seed_sk, seed_evk=seed_evk, seed_rk=seed_rk)
print("written")
