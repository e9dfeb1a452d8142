"""CPU tests of the circuit-level oracle (oracle/linformer_sim.py): the slot simulator against the independent
matrix-level float model, against the committed fixture, and its Chebyshev routine against numpy's own interpolation."""
import math
import os

import numpy as np
import pytest

from fhe_linformer_b200 import synth
from oracle import linformer_sim as ls

GOLD = os.path.join(os.path.dirname(__file__), "golden", "linformer_sim_s129.npz")


@pytest.mark.parametrize("S", [129, 200, 256])
def test_slot_simulator_matches_float_model(S):
    model = synth.make_model(n_classes=8)
    sample = synth.make_sample(model, S - 1, seed=S)
    z = ls.sim_forward(model, sample)
    f = ls.float_forward(model, sample)
    assert np.abs(z - f["logits"]).max() < 1e-12
    assert (z[8:] == 0).all()                      # classes 8..19 have zero weights
    # every approximation is used inside its fitting interval
    assert f["pre_gelu_max"] < 1 and f["pre_tanh_max"] < 1 and 0 < f["exp_sum"] <= 128 and f["h1_max"] < 1 and f["gelu_max"] < 1


def test_golden_fixture():
    g = np.load(GOLD)
    model = synth.make_model(n_classes=8)
    sample = synth.make_sample(model, 128, seed=20261018 + 1)
    cp = {}
    z = ls.sim_forward(model, sample, cp)
    assert np.abs(z - g["logits"]).max() < 1e-12
    for k in g.files:
        if k.startswith("cp_"):
            assert np.abs(cp[k[3:]][::97] - g[k]).max() < 1e-12, k


def test_chebyshev_restatement_is_interpolation():
    f = lambda x: math.tanh(3 * x) + 0.2 * x
    c = ls.chebyshev_coefficients(f, -2.0, 5.0, 63)
    # numpy's interpolant at the same (first-kind) nodes, mapped to [-2, 5]
    ref = np.polynomial.chebyshev.chebinterpolate(lambda u: np.tanh(3 * (3.5 * u + 1.5)) + 0.2 * (3.5 * u + 1.5), 63)
    mine = c.copy(); mine[0] *= 0.5
    assert np.abs(mine - ref).max() < 1e-12
    xs = np.linspace(-2, 5, 101)
    assert np.abs(ls.chebyshev_eval(c, -2, 5, xs) - np.array([f(x) for x in xs])).max() < 5e-4   # degree 63, pole at +-i pi/6


def test_layout_roundtrips():
    s = ls.SlotSim()
    rng = np.random.default_rng(0)
    rows = [rng.uniform(-1, 1, 128) for _ in range(5)]
    # CR-layout vectors (entry j at slot 128 j) -> wrapUpExpanded -> unwrapExpanded gives back Expanded vectors
    cr = [np.zeros(s.n) for _ in rows]
    for v, r in zip(cr, rows):
        v[::128] = r
    back = s.unwrapExpanded(s.wrapUpExpanded(cr), 5)
    for b, r in zip(back, rows):
        assert np.abs(b - s.expanded(r)).max() < 1e-15
    # containers: 512-wide hidden rows packed 32 per ciphertext and unpacked as four Repeated 128-vectors
    hid = [np.concatenate([rng.uniform(-1, 1, 512), np.zeros(s.n - 512)]) for _ in range(3)]
    quads = s.unwrapRepeatedLarge(s.generate_containers(hid), 3)
    for q, h in zip(quads, hid):
        for b in range(4):
            assert np.abs(q[b] - s.repeated(h[128 * b:128 * (b + 1)])).max() < 1e-15


def test_synthetic_files_roundtrip(tmp_path):
    model = synth.make_model(n_classes=5)
    sample = synth.make_sample(model, 130, seed=3)
    d = synth.write_files(str(tmp_path), model, sample)
    got = np.loadtxt(os.path.join(d["weights"], synth.LAYER + "selfAttn_WO_weight.txt"), delimiter=",")
    assert np.array_equal(got, model["WO"])
    assert len(os.listdir(d["tokens"])) == 130 and len(os.listdir(d["input"])) == 64
    assert np.array_equal(np.loadtxt(os.path.join(d["input"], "XF_31.txt"), delimiter=","), sample["XF"][31])
