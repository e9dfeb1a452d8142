"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the baby-step/giant-step diagonal matrix-vector product that OpenFHE's
EvalLinearTransform evaluates inside EvalBootstrap (reference call site FHEController.cpp:445; CoeffsToSlots / SlotsToCoeffs)
and that fhe_linformer_b200/csrc/lintrans.cpp plans and runs on ciphertexts.  Published algorithm (Halevi-Shoup diagonals with
the BSGS split of Cheon-Han-Hhan / Bossuat et al.); OpenFHE itself is absent from /root/reference, so parity is unpinned against it:

    (M v)[p] = sum_d diag_d[p] * v[(p + d) mod n],   d = g * (n1 * j + i - off)
    M v      = sum_j Rot_{G_j}( sum_i P_{j,i} * Rot_{g i}(v) ),   G_j = g * (n1 * j - off),   P_{j,i} = Rot_{-G_j}(diag_d)

Rot_k(v)[p] = v[(p + k) mod n] is EvalRotate(v, k) on the slots.  plan() mirrors Scheme::lintrans_plan (stride = gcd of the
shifts, n1 = the power of two >= sqrt(span), rotating giant steps first, the non-rotating one last, empty giant steps dropped);
apply() evaluates the right-hand side with numpy, one np.roll per rotation."""
import math

import numpy as np

K_BSGS_MAX = 16


def rot(v, k):
    return np.roll(v, -k)


def plan(diags, n, max_baby=0):
    """diags: {shift: vector[n]}.  Returns a dict with g, n1, n2, off, giant (rotation per giant step), P ([j][i] pre-rotated
    diagonals or None), mask and the rotation amounts the evaluation needs."""
    by_shift = {}
    g = 0
    for d, vec in diags.items():
        assert len(vec) == n
        d = d % n
        if d > n // 2:
            d -= n
        by_shift[d] = np.asarray(vec, np.complex128)
        g = math.gcd(g, abs(d))
    g = g or 1
    lo, hi = min(by_shift) // g, max(by_shift) // g
    off = -min(lo, 0)
    cnt = max(hi, 0) + off + 1
    n1 = 1
    while n1 * n1 < cnt:
        n1 <<= 1
    if max_baby > 0:
        n1 = min(n1, max_baby)
    n1 = min(n1, K_BSGS_MAX)
    n2 = (cnt + n1 - 1) // n1
    assert n2 <= K_BSGS_MAX
    order = [j for j in range(n2) if n1 * j != off] + [j for j in range(n2) if n1 * j == off]
    giant, P, mask = [], [], []
    for j in order:
        G = g * (n1 * j - off)
        row, m = [None] * n1, 0
        for i in range(n1):
            d = g * (n1 * j + i - off)
            if d in by_shift:
                row[i] = rot(by_shift[d], -G)
                m |= 1 << i
        if m:
            giant.append(G); P.append(row); mask.append(m)
    used = 0
    for m in mask:
        used |= m
    rotations = sorted({g * i for i in range(1, n1) if (used >> i) & 1} | {G for G in giant if G})
    return {"g": g, "n1": n1, "n2": len(giant), "off": off, "giant": giant, "P": P, "mask": mask, "rotations": rotations,
            "ndiag": sum(bin(m).count("1") for m in mask)}


def apply(pl, v):
    v = np.asarray(v, np.complex128)
    baby = [rot(v, pl["g"] * i) for i in range(pl["n1"])]
    acc = np.zeros_like(v)
    for G, row in zip(pl["giant"], pl["P"]):
        inner = np.zeros_like(v)
        for i, p in enumerate(row):
            if p is not None:
                inner += p * baby[i]
        acc += rot(inner, G)
    return acc


def matvec(diags, v):
    out = np.zeros(len(v), np.complex128)
    for d, vec in diags.items():
        out += np.asarray(vec) * rot(np.asarray(v, np.complex128), d)
    return out
