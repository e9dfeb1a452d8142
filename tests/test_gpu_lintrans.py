"""BSGS diagonal ciphertext x plaintext matrix product (fl_lt_*, csrc/lintrans.cpp, Engine::linear_transform) on the GPU.

Checked three ways: against the numpy matrix-vector product (the oracle of the operation), against the same plan evaluated as
separate EvalRotate / EvalMult / EvalAdd calls (fl_lt_apply_plain), and -- because both paths end in exact ModDowns of the same
integers up to key-switch rounding -- the two decryptions must agree far below the message precision."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from fhe_linformer_b200 import CKKS
    c = CKKS(logN=12, L=8, dnum=3, sparse_h=64)
    c.keygen(11)
    c.gen_mult_key()
    return c


def matvec(diags, v):
    n = len(v)
    out = np.zeros(n, np.complex128)
    for d, dv in diags.items():
        out += np.asarray(dv) * np.roll(v, -d)          # v[(p + d) mod n]
    return out


def run(c, diags, n, level=0, batch=1, max_baby=0, tol=2e-6):
    rng = np.random.default_rng(len(diags) * 7 + n)
    lt = c.linear_transform(diags, slots=n, level=level, max_baby=max_baby)
    from oracle import bsgs
    pl = bsgs.plan(diags, n, max_baby)                  # the engine plans the same split as the restatement
    assert lt.shape == {"n1": pl["n1"], "n2": pl["n2"], "stride": pl["g"], "diagonals": pl["ndiag"]}
    assert sorted(lt.rotations()) == pl["rotations"]
    c.gen_rot_keys(lt.rotations())
    vs = [rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n) for _ in range(batch)]
    cts = [c.encrypt(v, level=level, slots=n) for v in vs]
    outs = c.unpack(lt.apply(c.pack(cts))) if batch > 1 else [lt.apply(cts[0])]
    for v, ct, o in zip(vs, cts, outs):
        ref = matvec(diags, v)
        got = c.decrypt(o, slots=n, complex_out=True)
        assert o.deg == 2 and o.level == level
        assert np.abs(got - ref).max() < tol * max(1.0, np.abs(ref).max())
        plain = c.decrypt(lt.apply_plain(ct), slots=n, complex_out=True)
        assert np.abs(got - plain).max() < tol
    return lt


def test_dense_band(ctx):
    n = ctx.N // 2
    rng = np.random.default_rng(1)
    diags = {d: rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n) for d in range(-9, 12)}     # 21 diagonals: 8 baby x 3 giant steps
    lt = run(ctx, diags, n)
    s = lt.shape
    assert s["diagonals"] == 21 and s["n1"] * s["n2"] >= 21 and s["stride"] == 1


def test_wide_band_sixteen_baby_steps(ctx):
    """More than 64 diagonal positions: 16 baby steps (the 16-wide instantiation of the inner kernel) and 6 giant steps."""
    n = ctx.N // 2
    rng = np.random.default_rng(5)
    diags = {d: rng.uniform(-1, 1, n) for d in range(-40, 50) if d % 7}
    lt = run(ctx, diags, n, level=1)
    assert lt.shape["n1"] == 16 and lt.shape["n2"] == 6


def test_strided_sparse_diagonals_and_levels(ctx):
    n = ctx.N // 2
    rng = np.random.default_rng(2)
    diags = {d * 16: rng.uniform(-1, 1, n) for d in (-5, -3, 0, 1, 2, 7, 11)}                    # stride 16, holes in both steps
    lt = run(ctx, diags, n, level=2)
    assert lt.shape["stride"] == 16


def test_identity_only_and_single_rotation(ctx):
    n = ctx.N // 2
    run(ctx, {0: np.full(n, 0.5)}, n)
    run(ctx, {3: np.linspace(-1, 1, n)}, n)
    run(ctx, {-1: np.ones(n), 0: np.ones(n), 1: np.ones(n)}, n, max_baby=1)                       # giant steps only


def test_sparse_slots_and_batch(ctx):
    n = 256                                                                                      # sparse packing: n < N/2
    rng = np.random.default_rng(3)
    diags = {d: rng.uniform(-1, 1, n) for d in range(0, 40)}
    run(ctx, diags, n, batch=3)


def test_matrix_product_by_diagonals(ctx):
    """A dense d x d weight matrix applied to a vector replicated across the slots: the packed linear layer of the north star."""
    n, d = ctx.N // 2, 32
    rng = np.random.default_rng(4)
    W = rng.uniform(-1, 1, (d, d)) / np.sqrt(d)
    diags = {k: np.tile(np.array([W[p, (p + k) % d] for p in range(d)]), n // d) for k in range(d)}
    lt = ctx.linear_transform(diags, slots=n, level=0)
    ctx.gen_rot_keys(lt.rotations())
    x = rng.uniform(-1, 1, d)
    ct = ctx.encrypt(np.tile(x, n // d), level=0, slots=n)
    got = ctx.decrypt(lt.apply(ct), slots=n)
    assert np.abs(got[:d] - W @ x).max() < 2e-6 and np.abs(got[d:2 * d] - W @ x).max() < 2e-6


def test_missing_key_is_an_error(ctx):
    from fhe_linformer_b200 import CKKS
    c = CKKS(logN=11, L=4, dnum=2, sparse_h=32)
    c.keygen(3)
    n = c.N // 2
    lt = c.linear_transform({0: np.ones(n), 5: np.ones(n)}, slots=n, level=0)
    with pytest.raises(RuntimeError, match="no evaluation key for rotation 5"):
        lt.apply(c.encrypt(np.ones(n), level=0, slots=n))
