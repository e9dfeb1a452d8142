"""Scheme-level parity: keys, encoding and seeded encryption of the CUDA engine are
limb-for-limb those of the CPU restatement (oracle/ckks_oracle.c), at a small ring and at the reference ring
(FHEController.cpp:6-35: N = 2^15, 28 Q limbs, dnum 4); a rotate-and-add ladder without hoisting equals the sequential
ladder of FHEController.cpp:829-837 limb for limb.  Bar: bit-exact (integer work)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RINGS = [dict(logN=12, L=14, dnum=3), dict(logN=15, L=28, dnum=4)]


@pytest.fixture(scope="module", params=RINGS, ids=lambda p: "logN%d_L%d_d%d" % (p["logN"], p["L"], p["dnum"]))
def ctx(request):
    from fhe_linformer_b200 import CKKS
    from oracle.oracle import Oracle
    o = Oracle(**request.param)
    c = CKKS(sparse_h=64, **request.param)
    seed = 42
    c.keygen(seed)            # the seeded TEST entry (fl_keygen_seeded): every stream derives from `seed`
    c.gen_mult_key()
    c.gen_rot_keys([1, -1, 2, 4])
    c.gen_conj_key()
    sk = o.gen_sk(seed, h=64)
    pk = o.gen_pk(seed + 1, sk)
    yield o, c, seed, sk, pk
    c.close()


def test_key_pair_bit_exact(ctx):
    o, c, seed, sk, pk = ctx
    assert (c.export_sk() == sk).all(), "secret key limbs differ from the oracle"
    assert (c.export_pk() == pk).all(), "public key limbs differ from the oracle"


def test_relinearisation_key_bit_exact(ctx):
    o, c, seed, sk, pk = ctx
    assert (c.export_evk(0) == o.gen_relin_key(seed + 2, sk)).all()


@pytest.mark.parametrize("k", [1, -1, 4])
def test_rotation_keys_bit_exact(ctx, k):
    o, c, seed, sk, pk = ctx
    g = o.galois(k)
    assert (c.export_evk(g) == o.gen_galois_key(seed + 1000 + g, sk, g)).all()


def test_conjugation_key_bit_exact(ctx):
    o, c, seed, sk, pk = ctx
    g = o.galois_conj()
    assert (c.export_evk(g) == o.gen_galois_key(seed + 1000 + g, sk, g)).all()


def test_encode_bit_exact_two_levels(ctx):
    o, c, seed, sk, pk = ctx
    n = o.N // 2
    rng = np.random.default_rng(0)
    v, w = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    assert (c.encode(v, level=0).export()[0] == o.encode(v, o.sf[0], o.L)).all(), "level 0"
    assert (c.encode(v + 1j * w, level=3, slots=n).export()[0] == o.encode(v + 1j * w, o.sf[3], o.L - 3)).all(), "level 3, complex"
    # sparse packing (fewer slots than N/2) and edge vectors
    assert (c.encode(v[: n // 4], level=1, slots=n // 4).export()[0] == o.encode(v[: n // 4], o.sf[1], o.L - 1, slots=n // 4)).all(), "sparse packing"
    z = np.zeros(n); e1 = np.zeros(n); e1[0] = 1.0
    for name, x in (("zeros", z), ("unit", e1), ("ones", np.ones(n))):
        assert (c.encode(x, level=0).export()[0] == o.encode(x, o.sf[0], o.L)).all(), name


def test_seeded_encrypt_bit_exact(ctx):
    o, c, seed, sk, pk = ctx
    n = o.N // 2
    v = np.random.default_rng(1).uniform(-1, 1, n)
    pt = c.encode(v, level=0)
    ct = c.encrypt(pt, seed=7)
    assert (ct.export() == o.encrypt(7, o.encode(v, o.sf[0], o.L), pk)).all()
    # a lower level: the ciphertext uses the first L - 2 limbs of the public key (the oracle takes the full key)
    pt2 = c.encode(v, level=2)
    ref = o.encrypt(9, o.encode(v, o.sf[2], o.L - 2), pk)
    assert (c.encrypt(pt2, seed=9).export() == ref).all()
    assert np.abs(c.decrypt(ct) - v).max() < 1e-7


def test_unseeded_encryptions_never_repeat(ctx):
    """fl_encrypt draws from ChaCha20 streams keyed from the operating system: two encryptions of one plaintext differ in
    every limb, and both decrypt."""
    o, c, seed, sk, pk = ctx
    v = np.random.default_rng(2).uniform(-1, 1, o.N // 2)
    pt = c.encode(v, level=0)
    a, b = c.encrypt(pt).export(), c.encrypt(pt).export()
    assert (a != b).mean() > 0.99
    assert np.abs(c.decrypt(c.encrypt(pt)) - v).max() < 1e-7


def test_sequential_ladder_limb_exact(ctx):
    """rotsum with only the doubling keys present (no hoisted groups) is r <- r + EvalRotate(r, stride 2^i) step by step:
    the same limbs as the oracle's sequential ladder (FHEController.cpp:829-837)."""
    o, c, seed, sk, pk = ctx
    l = o.L - 1
    rng = np.random.default_rng(3)
    q = [int(x) for x in o.moduli]
    ct = np.stack([np.stack([rng.integers(0, q[m], o.N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
    steps, stride = 3, 1                                   # keys 1, 2, 4 exist; 3, 5, 6, 7 do not, so no group is hoisted
    got = c.rotsum(c.import_elem(ct, 1, float(o.sf[1]), o.N // 2), steps, stride).export()
    ref = ct.copy()
    for i in range(steps):
        g = o.galois(stride << i)
        rot = o.rotate(ref, g, o.gen_galois_key(seed + 1000 + g, sk, g))
        ref = np.stack([o.add(ref[0], rot[0], list(range(l))), o.add(ref[1], rot[1], list(range(l)))])
    assert (got == ref).all()


def test_unseeded_keys_are_fresh_and_work():
    """fl_keygen(ctx, 0): two contexts never share a key, nothing derives from a seed, encryption round-trips."""
    from fhe_linformer_b200 import CKKS
    P = dict(logN=11, L=4, dnum=2)
    a, b = CKKS(sparse_h=32, **P), CKKS(sparse_h=32, **P)
    a.keygen(); b.keygen()
    assert (a.export_sk() != b.export_sk()).any() and (a.export_pk() != b.export_pk()).mean() > 0.99
    sk = a.export_sk()
    a.keygen()   # a second call draws new keys
    assert (a.export_sk() != sk).any()
    a.gen_mult_key(); a.gen_rot_keys([1])
    v = np.random.default_rng(4).uniform(-1, 1, a.N // 2)
    ct = a.encrypt(v)
    assert np.abs(a.decrypt(a.rotate(a.mult(ct, ct), 1)) - np.roll(v * v, -1)).max() < 1e-6
    a.close(); b.close()


def test_first_use_of_a_slot_count_from_two_threads():
    """Two controllers, each on its own host thread, hit encode / decode for slot counts no one in this process has used yet:
    the special-FFT table cache is shared by all contexts and must tolerate concurrent first use (no pre-warming here)."""
    import threading
    from fhe_linformer_b200 import CKKS
    P = dict(logN=11, L=3, dnum=2)
    ctxs = [CKKS(sparse_h=32, **P) for _ in range(2)]
    for c in ctxs: c.keygen()
    errs, fails = {}, []
    barrier = threading.Barrier(2)

    def work(i, c):
        try:
            rng = np.random.default_rng(100 + i)
            barrier.wait()
            worst = 0.0
            for slots in (8, 16, 64, 128, 8, 16):            # unusual sizes: cold in the cache on first touch
                v = rng.uniform(-1, 1, slots)
                worst = max(worst, float(np.abs(c.decrypt(c.encrypt(v, slots=slots)) - v).max()))
            errs[i] = worst
        except Exception as e:   # noqa: BLE001
            fails.append(repr(e))

    th = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate(ctxs)]
    for t in th: t.start()
    for t in th: t.join()
    assert not fails, fails
    assert len(errs) == 2 and max(errs.values()) < 1e-6
    for c in ctxs: c.close()


def test_veneer_sequential_ladder_limb_exact(tmp_path):
    """Through the FHEController veneer with batch_rows = false and hoist_ladders = false, rotsum(c, 8, 128) is the reference's loop
    r = r + EvalRotate(r, 128 * 2^i) (FHEController.cpp:829-837), and its limbs are the CPU restatement's, at the reference ring."""
    from fhe_linformer_b200 import host
    from oracle.oracle import Oracle
    seed = 4242
    fc = host.FHEController(root=str(tmp_path), key_seed=seed, batch_rows=0, hoist_ladders=0).generate(rotations=[128, 256, 512])
    c = fc.ckks
    o = Oracle(logN=15, L=28, dnum=4)
    sk = o.gen_sk(seed, h=192)                             # SPARSE_TERNARY as the controller generates it
    assert (c.export_sk() == sk).all()
    l = 27
    rng = np.random.default_rng(9)
    q = [int(x) for x in o.moduli]
    ct = np.stack([np.stack([rng.integers(0, q[m], o.N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
    got = fc.invoke("rotsum", [c.import_elem(ct, 1, float(o.sf[1]), o.N // 2)], ints=[8, 128])[0].export()
    ref = ct.copy()
    for i in range(3):
        g = o.galois(128 << i)
        rot = o.rotate(ref, g, o.gen_galois_key(seed + 1000 + g, sk, g))
        ref = np.stack([o.add(ref[0], rot[0], list(range(l))), o.add(ref[1], rot[1], list(range(l)))])
    assert (got == ref).all()
    # a rotation index without a key is an error, as from OpenFHE's EvalRotate -- the veneer no longer makes keys on demand
    with pytest.raises(RuntimeError, match="no evaluation key"):
        fc.invoke("rotate", [c.import_elem(ct, 1, float(o.sf[1]), o.N // 2)], ints=[11111])   # not a bootstrap index either
    fc.close()


def test_batched_encryption(ctx):
    """fl_encrypt_many: every ciphertext of the batch decrypts to its own plaintext and carries its own randomness."""
    o, c, seed, sk, pk = ctx
    n = o.N // 2
    rng = np.random.default_rng(6)
    vs = [rng.uniform(-1, 1, n) for _ in range(5)]
    cts = c.encrypt_many([c.encode(v, level=1) for v in vs])
    assert len(cts) == 5
    for v, ct in zip(vs, cts):
        assert ct.level == 1 and np.abs(c.decrypt(ct) - v).max() < 1e-7
    same = c.encrypt_many([c.encode(vs[0], level=0)] * 3)            # one plaintext three times: three different ciphertexts
    a, b, d = (x.export() for x in same)
    assert (a != b).mean() > 0.99 and (a != d).mean() > 0.99 and (b != d).mean() > 0.99
    with pytest.raises(RuntimeError):
        c.encrypt_many([c.encode(vs[0], level=0), c.encode(vs[1], level=2)])   # mixed levels
    # fl_encode_many: the batched plaintext holds, element by element, exactly the limbs of the single encodings (short
    # rows are zero-padded like fl_encode's), and encrypts as one operand
    rows = np.stack([v[: n // 2] for v in vs])
    many = c.encode_many(rows, level=2)
    parts = c.unpack(many)
    assert len(parts) == 5
    for r, pt in zip(rows, parts):
        assert (pt.export() == c.encode(r, level=2).export()).all()
    for r, ct in zip(rows, c.encrypt_many(many)):
        assert np.abs(c.decrypt(ct)[: n // 2] - r).max() < 1e-7 and np.abs(c.decrypt(ct)[n // 2:]).max() < 1e-7
    # fl_encrypt_values_many: encode + encrypt in one pass (message and e0 share a transform); same decryptions, own randomness
    direct = c.encrypt_values_many(rows, level=2)
    assert len(direct) == 5
    for r, ct in zip(rows, direct):
        assert ct.level == 2 and np.abs(c.decrypt(ct)[: n // 2] - r).max() < 1e-7 and np.abs(c.decrypt(ct)[n // 2:]).max() < 1e-7
    twice = c.encrypt_values_many(np.stack([rows[0], rows[0]]), level=0)
    assert (twice[0].export() != twice[1].export()).mean() > 0.99
    got = c.encrypt_many(parts[1:4])                                            # slices of a batched plaintext: the contiguous path
    for r, ct in zip(rows[1:4], got):
        assert np.abs(c.decrypt(ct)[: n // 2] - r).max() < 1e-7
