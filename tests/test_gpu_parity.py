"""Parity tests proper: every CUDA primitive, called through the C-ABI (include/fl_ckks.h), against the CPU
oracle on the same seeded inputs.  Bar: bit-exact limbs (integer work)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rnd(o, rng, midx):
    q = [int(x) for x in o.moduli]
    return np.stack([rng.integers(0, q[m], o.N, dtype=np.uint64) for m in midx])


def rnd_ct(o, rng, l):
    return np.stack([rnd(o, rng, range(l)), rnd(o, rng, range(l))])


@pytest.fixture(scope="module", params=[dict(logN=10, L=6, dnum=3), dict(logN=12, L=9, dnum=3), dict(logN=13, L=5, dnum=5)],
                ids=lambda p: "logN%d_L%d_d%d" % (p["logN"], p["L"], p["dnum"]))
def pair(request):
    from fhe_linformer_b200 import Engine
    from oracle.oracle import Oracle
    o = Oracle(**request.param)
    e = Engine(device=0, sparse_h=64, **request.param)
    yield o, e
    e.close()


def test_context_tables_agree(pair):
    o, e = pair
    assert (o.moduli == e.moduli).all() and (o.roots == e.roots).all() and (o.sf == e.sf).all()
    assert (o.L, o.K, o.alpha, o.dnum) == (e.L, e.K, e.alpha, e.dnum)
    for k in [1, -1, 3, -64]:
        assert o.galois(k) == e.galois(k)


def test_ntt_intt_all_moduli(pair):
    o, e = pair
    rng = np.random.default_rng(1)
    allm = list(range(o.L + o.K))
    a = rnd(o, rng, allm)
    d = e.to_dev(a)
    assert (e.ntt(d, allm).download() == o.ntt(a, allm)).all()
    assert (e.intt(d, allm).download() == a).all()
    # edge inputs: zeros, q-1 everywhere, a single one
    for special in [np.zeros_like(a), (o.moduli[allm][:, None] - 1) * np.ones_like(a), np.eye(1, o.N, 0, dtype=np.uint64).repeat(len(allm), 0)]:
        special = np.ascontiguousarray(special, np.uint64)
        assert (e.ntt(e.to_dev(special), allm).download() == o.ntt(special, allm)).all()
    assert (e.host_ntt(a[: o.L]) == o.ntt(a[: o.L])).all()


def test_elementwise_and_automorphism(pair):
    o, e = pair
    rng = np.random.default_rng(2)
    allm = list(range(o.L + o.K))
    a, b = rnd(o, rng, allm), rnd(o, rng, allm)
    da, db = e.to_dev(a), e.to_dev(b)
    for nm in ["add", "sub", "mul"]:
        assert (getattr(e, nm)(da, db, allm).download() == getattr(o, nm)(a, b, allm)).all(), nm
    for g in [o.galois(1), o.galois(-3), o.galois(o.N // 4), o.galois_conj()]:
        assert (e.automorph(da, g).download() == o.automorph_eval(a, g)).all()
    ct, pt = rnd_ct(o, rng, o.L), rnd(o, rng, range(o.L))
    assert (e.mul_plain(e.to_dev(ct), e.to_dev(pt)).download() == o.mul_plain(ct, pt)).all()


def test_rescale_modup_moddown_every_level(pair):
    o, e = pair
    rng = np.random.default_rng(3)
    for l in range(o.L, 0, -1):
        x = rnd(o, rng, range(l))
        dx = e.to_dev(x)
        if l >= 2:
            assert (e.rescale(dx).download() == o.rescale(x)).all(), ("rescale", l)
            ct = rnd_ct(o, rng, l)
            want = np.stack([o.rescale(ct[0]), o.rescale(ct[1])])
            assert (e.rescale(e.to_dev(ct)).download() == want).all()
        for dg in range((l + o.alpha - 1) // o.alpha):
            assert (e.modup(dx, dg).download() == o.modup(x, dg)).all(), ("modup", l, dg)
        ext = list(range(l)) + [o.L + k for k in range(o.K)]
        xe = rnd(o, rng, ext)
        assert (e.moddown(e.to_dev(xe)).download() == o.moddown(xe)).all(), ("moddown", l)


def test_keyswitch_rotate_multiply_every_level(pair):
    o, e = pair
    rng = np.random.default_rng(4)
    sk = o.gen_sk(7, h=64)
    g = o.galois(-1)
    evk, rk = o.gen_galois_key(11, sk, g), o.gen_relin_key(12, sk)
    devk, drk = e.to_dev(evk), e.to_dev(rk)
    for l in range(o.L, 0, -1):
        a, b = rnd_ct(o, rng, l), rnd_ct(o, rng, l)
        k0, k1 = o.keyswitch(a[1], evk)
        ks = e.keyswitch(e.to_dev(a[1]), devk).download()
        assert (ks[0] == k0).all() and (ks[1] == k1).all(), ("keyswitch", l)
        want = o.rotate(a, g, evk)
        assert (e.rotate(e.to_dev(a), g, devk).download() == want).all(), ("rotate", l)
        assert (e.host_rotate(a, g, devk) == want).all(), ("host_rotate", l)
        wantm = o.mul_relin(a, b, rk)
        assert (e.mul_relin(e.to_dev(a), e.to_dev(b), drk).download() == wantm).all(), ("mul_relin", l)
        assert (e.host_mul_relin(a, b, drk) == wantm).all()


def test_golden_fixture_on_gpu():
    """Committed vectors (no oracle involved on this path)."""
    from fhe_linformer_b200 import Engine
    from oracle.oracle import Oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_ring.npz"))
    e = Engine(device=0, logN=10, L=6, dnum=3, sparse_h=64)
    assert (g["moduli"] == e.moduli).all() and (g["roots"] == e.roots).all()
    assert (e.ntt(e.to_dev(g["poly_coeff"])).download() == g["poly_eval"]).all()
    o = Oracle(logN=10, L=6, dnum=3)   # only to regenerate the keys from their seeds
    sk = o.gen_sk(int(g["seed_sk"]), h=64)
    gal = o.galois(1)
    d_ct = e.to_dev(g["ct"])
    assert (e.rotate(d_ct, gal, e.to_dev(o.gen_galois_key(int(g["seed_evk"]), sk, gal))).download() == g["rotated"]).all()
    m = e.mul_relin(d_ct, d_ct, e.to_dev(o.gen_relin_key(int(g["seed_rk"]), sk)))
    assert (m.download() == g["mult"]).all()
    assert (e.rescale(m).download() == g["rescaled"]).all()
    e.close()


def test_reference_parameters_N32768():
    """Reference CKKS parameters (FHEController.cpp:6-35): N=2^15, 28+7 limbs, dnum=4, selected levels."""
    from fhe_linformer_b200 import Engine
    from oracle.oracle import Oracle
    o = Oracle(logN=15, L=28, dnum=4)
    e = Engine(device=0)
    assert (o.moduli == e.moduli).all() and (o.roots == e.roots).all()
    rng = np.random.default_rng(5)
    sk = o.gen_sk(1, h=192)
    g = o.galois(128)
    evk, rk = o.gen_galois_key(2, sk, g), o.gen_relin_key(3, sk)
    devk, drk = e.to_dev(evk), e.to_dev(rk)
    allm = list(range(35))
    a = rnd(o, rng, allm)
    assert (e.ntt(e.to_dev(a), allm).download() == o.ntt(a, allm)).all()
    for l in [28, 27, 15, 7, 1]:
        ct, ct2 = rnd_ct(o, rng, l), rnd_ct(o, rng, l)
        assert (e.rotate(e.to_dev(ct), g, devk).download() == o.rotate(ct, g, evk)).all(), l
        assert (e.mul_relin(e.to_dev(ct), e.to_dev(ct2), drk).download() == o.mul_relin(ct, ct2, rk)).all(), l
        if l > 1:
            assert (e.rescale(e.to_dev(ct)).download() == np.stack([o.rescale(ct[0]), o.rescale(ct[1])])).all()
    e.close()


def test_full_size_properties_N65536():
    """BASELINE config 2 sizes (N=2^16, full chain): size-independent properties instead of a full oracle sweep."""
    from fhe_linformer_b200 import Engine
    from oracle.oracle import Oracle
    o = Oracle(logN=16, L=28, dnum=4)
    e = Engine(device=0, logN=16)
    assert (o.moduli == e.moduli).all()
    rng = np.random.default_rng(6)
    allm = list(range(35))
    a = rnd(o, rng, allm)
    d = e.to_dev(a)
    assert (e.ntt(d, allm).download() == o.ntt(a, allm)).all()
    assert (e.intt(d, allm).download() == a).all()                      # INTT(NTT(x)) == x
    # linearity of the key switch: KS(a + b) == KS(a) + KS(b) limb-exactly is NOT guaranteed (approximate
    # base conversion), but rotate(+k) then rotate(-k) must decrypt back to the same slots.
    sk = o.gen_sk(1, h=192)
    pk = o.gen_pk(2, sk)
    n = o.N // 2
    v = rng.normal(size=n)
    ct = o.encrypt(3, o.encode(v, o.sf[0], 28), pk)
    gp, gm = o.galois(5), o.galois(-5)
    kp, km = e.to_dev(o.gen_galois_key(4, sk, gp)), e.to_dev(o.gen_galois_key(5, sk, gm))
    r1 = e.rotate(e.to_dev(ct), gp, kp)
    back = e.rotate(r1, gm, km).download()
    assert np.abs(o.decode(o.decrypt(r1.download(), sk), o.sf[0], n).real - np.roll(v, -5)).max() < 1e-6
    assert np.abs(o.decode(o.decrypt(back, sk), o.sf[0], n).real - v).max() < 1e-6
    # one full-size bit-exact rotation against the oracle
    evk = o.gen_galois_key(4, sk, gp)
    assert (r1.download() == o.rotate(ct, gp, evk)).all()
    e.close()


def test_batched_rotation_equals_single(pair):
    """A batch of ciphertexts through one batched key switch gives, per ciphertext, the oracle's limbs."""
    o, e = pair
    rng = np.random.default_rng(21)
    sk = o.gen_sk(3, h=64)
    g = o.galois(5)
    evk = o.gen_galois_key(9, sk, g)
    d_evk = e.to_dev(evk)
    for l in sorted({o.L, max(1, o.L - 2)}):
        cts = np.stack([rnd_ct(o, rng, l) for _ in range(5)])        # 5 is not a multiple of the 4-wide inner-product group
        got = e.rotate_batch(e.to_dev(cts), g, d_evk).download()
        for b in range(5):
            assert (got[b] == o.rotate(cts[b], g, evk)).all(), (l, b)
        assert (e.host_rotate_batch(cts, g, d_evk) == got).all()
        # overlapped calls (no wait inside the call, two staging slots used alternately, one sync at the end)
        outs = [np.zeros_like(cts) for _ in range(5)]
        ins = [np.roll(cts, i, axis=0).copy() for i in range(5)]
        for i in range(5):
            e.host_rotate_batch(ins[i], g, d_evk, out=outs[i], wait=False)
        e.sync()
        for i in range(5):
            assert (outs[i] == np.roll(got, i, axis=0)).all(), (l, i)


def test_batched_ntt_equals_single(pair):
    o, e = pair
    rng = np.random.default_rng(31)
    allm = list(range(o.L + o.K))
    a = np.stack([rnd(o, rng, allm) for _ in range(3)])
    d = e.to_dev(a)
    got = e.ntt_batch(d, allm).download()
    for b in range(3):
        assert (got[b] == o.ntt(a[b], allm)).all()
    assert (e.ntt_batch(d, allm, inverse=True).download() == a).all()


def test_radix16_transform_variant_is_bit_exact():
    """The alternative radix-16 head/tail kernels (FLK_NTT_RADIX16=1, logN >= 12) against the oracle, in a fresh process."""
    import subprocess, sys
    code = r'''
import numpy as np
from fhe_linformer_b200 import Engine
from oracle.oracle import Oracle
for P in (dict(logN=12, L=5, dnum=3), dict(logN=13, L=4, dnum=2), dict(logN=15, L=3, dnum=3), dict(logN=16, L=2, dnum=2)):
    o = Oracle(**P); e = Engine(device=0, sparse_h=64, **P)
    rng = np.random.default_rng(P["logN"])
    allm = list(range(o.L + o.K))
    a = np.stack([rng.integers(0, int(o.moduli[m]), o.N, dtype=np.uint64) for m in allm])
    d = e.to_dev(a)
    assert (e.ntt(d, allm).download() == o.ntt(a, allm)).all(), P
    assert (e.intt(d, allm).download() == a).all(), P
    e.close()
print("radix16 ok")
'''
    env = dict(os.environ, FLK_NTT_RADIX16="1", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "radix16 ok" in r.stdout, r.stderr[-2000:]
