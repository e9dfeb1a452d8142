"""The BSGS restatement (oracle/bsgs.py) against the direct diagonal product, on the CPU."""
import numpy as np
import pytest

from oracle import bsgs


@pytest.mark.parametrize("shifts,n", [
    (list(range(-9, 12)), 64),                       # dense band around zero
    ([16 * d for d in (-5, -3, 0, 1, 2, 7, 11)], 512),  # strided, holes in both steps
    ([0], 32), ([3], 32), ([-1, 0, 1], 32),
    (list(range(0, 40)), 256),
    ([d for d in range(-40, 50) if d % 7], 2048),   # 16 baby steps
])
def test_plan_evaluates_the_matrix(shifts, n):
    rng = np.random.default_rng(len(shifts) + n)
    diags = {d: rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n) for d in shifts}
    v = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    pl = bsgs.plan(diags, n)
    assert np.abs(bsgs.apply(pl, v) - bsgs.matvec(diags, v)).max() < 1e-12
    assert pl["ndiag"] == len(shifts) and pl["n1"] <= 16 and pl["n2"] <= 16
    assert all(G != 0 for G in pl["giant"][:-1])     # a non-rotating giant step, if any, comes last


def test_special_fft_stage_structure():
    """A collapsed CoeffsToSlots-like stage: 2^5 - 1 ... the merged butterflies of r levels have 2^(r+1) - 1 diagonals, stride 2^s."""
    n, stride, r = 1024, 8, 3
    shifts = [stride * i for i in range(-(2 ** r - 1), 2 ** r)]
    rng = np.random.default_rng(1)
    diags = {d: rng.uniform(-1, 1, n) for d in shifts}
    pl = bsgs.plan(diags, n)
    assert pl["g"] == stride and pl["n1"] == 4 and pl["n2"] == 4
    assert pl["rotations"] == [-56, -24, 8, 16, 24, 40]      # babies 8, 16, 24 and giants -56, -24, 8, 40: the key for 8 serves both
    v = rng.uniform(-1, 1, n)
    assert np.abs(bsgs.apply(pl, v) - bsgs.matvec(diags, v)).max() < 1e-12
