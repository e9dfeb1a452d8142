"""CPU oracle pinned by algebra and by an independent pure-Python big-int model (small rings only).

The reference holds no golden vectors for this path, so these identities plus the
committed fixtures in tests/golden/ are what pins the oracle ("parity unpinned" w.r.t. OpenFHE itself)."""
import numpy as np
import pytest


def rnd(o, rng, midx):
    q = [int(x) for x in o.moduli]
    return np.stack([rng.integers(0, q[m], o.N, dtype=np.uint64) for m in midx])


def crt(res, mods):
    """big-int CRT of residues res[i] mod mods[i]"""
    M = 1
    for m in mods:
        M *= m
    x = 0
    for r, m in zip(res, mods):
        Mi = M // m
        x += int(r) * Mi * pow(Mi, -1, m)
    return x % M, M


def test_prime_chain_reference_params(ref_oracle):
    o = ref_oracle
    q = [int(x) for x in o.moduli]
    assert (o.L, o.K, o.alpha, o.dnum) == (28, 7, 7, 4)          # SURVEY section 8 header
    assert len(set(q)) == 35
    for m in q:
        assert m % (2 * o.N) == 1
    assert q[0].bit_length() == 55 and all(51 <= x.bit_length() <= 53 for x in q[1:28]) and all(x.bit_length() == 60 for x in q[28:])
    # FLEXIBLEAUTO scaling factors stay close to 2^52 at every level (Appendix A.8)
    assert np.all(np.abs(np.log2(o.sf) - 52) < 0.01)


def test_roots_minimal(small_oracle):
    o = small_oracle
    for q, psi in zip(o.moduli[:3], o.roots[:3]):
        q, psi = int(q), int(psi)
        assert pow(psi, o.N, q) == q - 1
        g = psi
        best = min(pow(g, k, q) for k in range(1, 2 * o.N, 2))
        assert best == psi


def test_ntt_matches_direct_evaluation(small_oracle):
    o = small_oracle
    rng = np.random.default_rng(1)
    a = rnd(o, rng, range(2))
    A = o.ntt(a)
    br = lambda k: int(format(k, "0%db" % o.logN)[::-1], 2)
    for limb in range(2):
        q, psi = int(o.moduli[limb]), int(o.roots[limb])
        for k in [0, 1, 2, 77, o.N - 1]:
            e = pow(psi, 2 * br(k) + 1, q)
            v = sum(int(a[limb][j]) * pow(e, j, q) for j in range(o.N)) % q
            assert v == int(A[limb][k])
    assert (o.intt(A) == a).all()


def test_negacyclic_convolution(small_oracle):
    o = small_oracle
    rng = np.random.default_rng(2)
    a, b = rnd(o, rng, [0, o.L]), rnd(o, rng, [0, o.L])
    midx = [0, o.L]
    c = o.intt(o.mul(o.ntt(a, midx), o.ntt(b, midx), midx), midx)
    N = o.N
    for limb, m in enumerate(midx):
        q = int(o.moduli[m])
        aa, bb = [int(x) for x in a[limb]], [int(x) for x in b[limb]]
        ref = [0] * N
        for i in range(0, N, 37):       # sparse rows keep the schoolbook product quick
            pass
        # full schoolbook on a sparsified a
        sp = [(i, aa[i]) for i in range(0, N, 29)]
        a2 = np.zeros(N, np.uint64)
        for i, v in sp:
            a2[i] = v
        c2 = o.intt(o.mul(o.ntt(a2[None, :], [m]), o.ntt(b[limb][None, :], [m]), [m]), [m])[0]
        for i, v in sp:
            for j in range(N):
                k = i + j
                if k < N:
                    ref[k] += v * bb[j]
                else:
                    ref[k - N] -= v * bb[j]
        assert [r % q for r in ref] == [int(x) for x in c2]
    assert c.shape == a.shape


def test_automorphism_eval_equals_coeff(small_oracle):
    o = small_oracle
    rng = np.random.default_rng(3)
    a = rnd(o, rng, range(3))
    for k in [1, -1, 5, 64, -64]:
        g = o.galois(k)
        assert (o.automorph_eval(o.ntt(a), g) == o.ntt(o.automorph_coeff(a, g))).all()
    g = o.galois_conj()
    assert (o.automorph_eval(o.ntt(a), g) == o.ntt(o.automorph_coeff(a, g))).all()
    assert o.galois(1) == 5 and (o.galois(-1) * 5) % (2 * o.N) == 1


def test_rescale_matches_bigint(small_oracle):
    o = small_oracle
    rng = np.random.default_rng(4)
    for l in [6, 3, 2]:
        x = rnd(o, rng, range(l))
        out = o.intt(o.rescale(o.ntt(x)))
        mods = [int(m) for m in o.moduli[:l]]
        ql = mods[-1]
        for j in [0, 1, 500, o.N - 1]:
            X, _ = crt([x[i][j] for i in range(l)], mods)
            last = int(x[l - 1][j])
            cent = last - ql if last > ql // 2 else last
            y = (X - cent) // ql                      # exact: X - cent is divisible by q_last
            assert (X - cent) % ql == 0
            for i in range(l - 1):
                assert y % mods[i] == int(out[i][j])


def test_modup_moddown_match_bigint(small_oracle):
    o = small_oracle
    rng = np.random.default_rng(5)
    q = [int(m) for m in o.moduli]
    for l in [6, 5, 3]:
        x = rnd(o, rng, range(l))
        X = o.ntt(x)
        for d in range((l + o.alpha - 1) // o.alpha):
            lo, hi = d * o.alpha, min((d + 1) * o.alpha, l)
            up = o.modup(X, d)
            ext = list(range(l)) + [o.L + k for k in range(o.K)]
            upc = o.intt(up, ext)
            src = q[lo:hi]
            Qd = int(np.prod([1])) * 1
            for s in src:
                Qd *= s
            for j in [0, 3, o.N - 1]:
                y = [(int(x[lo + i][j]) * pow(Qd // src[i], -1, src[i])) % src[i] for i in range(hi - lo)]
                for t, m in enumerate(ext):
                    want = sum(y[i] * ((Qd // src[i]) % q[m]) for i in range(hi - lo)) % q[m]
                    if lo <= t < hi:
                        want = int(x[t][j])
                    assert want == int(upc[t][j])
        # ModDown: (c - conv(c_P)) / P on the Q limbs
        ext = list(range(l)) + [o.L + k for k in range(o.K)]
        z = rnd(o, rng, ext)
        out = o.intt(o.moddown(o.ntt(z, ext)))
        pm = q[o.L:]
        Pp = 1
        for p in pm:
            Pp *= p
        for j in [0, 9, o.N - 1]:
            y = [(int(z[l + k][j]) * pow(Pp // pm[k], -1, pm[k])) % pm[k] for k in range(o.K)]
            for i in range(l):
                conv = sum(y[k] * ((Pp // pm[k]) % q[i]) for k in range(o.K)) % q[i]
                want = ((int(z[i][j]) - conv) * pow(Pp, -1, q[i])) % q[i]
                assert want == int(out[i][j])


def test_encode_decode_and_slot_rotation(small_oracle):
    o = small_oracle
    rng = np.random.default_rng(6)
    n = o.N // 2
    v = rng.normal(size=n) + 1j * rng.normal(size=n)
    pt = o.encode(v, o.sf[0], 3)
    assert np.abs(o.decode(pt, o.sf[0], n) - v).max() < 1e-9
    for k in [1, 7, -3]:
        rot = o.automorph_eval(pt, o.galois(k))
        assert np.abs(o.decode(rot, o.sf[0], n) - np.roll(v, -k)).max() < 1e-9
    conj = o.automorph_eval(pt, o.galois_conj())
    assert np.abs(o.decode(conj, o.sf[0], n) - v.conj()).max() < 1e-9
    # sparse packing
    w = rng.normal(size=64)
    assert np.abs(o.decode(o.encode(w, o.sf[0], 2, slots=64), o.sf[0], 64) - w).max() < 1e-9


def test_encrypt_keyswitch_rotate_mult_decrypt(small_oracle):
    o = small_oracle
    rng = np.random.default_rng(7)
    n = o.N // 2
    sk = o.gen_sk(7, h=64)
    pk = o.gen_pk(8, sk)
    v = rng.normal(size=n) + 1j * rng.normal(size=n)
    ct = o.encrypt(9, o.encode(v, o.sf[0], o.L), pk)
    assert np.abs(o.decode(o.decrypt(ct, sk), o.sf[0], n) - v).max() < 1e-8
    for l in [6, 4, 3, 1]:
        for k in [1, -2]:
            g = o.galois(k)
            evk = o.gen_galois_key(100 + k, sk, g)
            r = o.rotate(ct[:, :l].copy(), g, evk)
            assert np.abs(o.decode(o.decrypt(r, sk), o.sf[0], n) - np.roll(v, -k)).max() < 1e-7
    rk = o.gen_relin_key(77, sk)
    m = o.mul_relin(ct, ct, rk)
    assert np.abs(o.decode(o.decrypt(m, sk), o.sf[0] ** 2, n) - v * v).max() < 1e-7
    m2 = np.stack([o.rescale(m[0]), o.rescale(m[1])])
    assert np.abs(o.decode(o.decrypt(m2, sk), o.sf[0] ** 2 / float(o.moduli[o.L - 1]), n) - v * v).max() < 1e-7
    # sk has the requested Hamming weight and the Gaussian looks like sigma = 3.19
    s = o.sample_sparse(1, 64)
    assert np.count_nonzero(s) == 64 and set(np.unique(s)) <= {-1, 0, 1}
    e = o.sample_gauss(2).astype(np.float64)
    assert abs(e.std() - 3.19) < 0.3 and abs(e.mean()) < 0.3


def test_golden_fixture(small_oracle):
    """Committed fixture (tests/golden/make_golden.py): pins the oracle's outputs across refactors."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_ring.npz"))
    o = small_oracle
    assert (g["moduli"] == o.moduli).all() and (g["roots"] == o.roots).all()
    ct = g["ct"]
    sk = o.gen_sk(int(g["seed_sk"]), h=64)
    gal = o.galois(1)
    assert (o.ntt(g["poly_coeff"]) == g["poly_eval"]).all()
    assert (o.rotate(ct, gal, o.gen_galois_key(int(g["seed_evk"]), sk, gal)) == g["rotated"]).all()
    assert (o.mul_relin(ct, ct, o.gen_relin_key(int(g["seed_rk"]), sk)) == g["mult"]).all()
    assert (np.stack([o.rescale(g["mult"][0]), o.rescale(g["mult"][1])]) == g["rescaled"]).all()
