"""Synthetic weights and inputs for the encrypted Linformer forward, written as the text files the reference reads.

The reference ships no weights, inputs or embeddings (all git-ignored / missing blobs), so every
workload here is synthetic: same shapes, same file names (src/main.cpp:161-455), same text format (`%.18e`, comma
separated, src/python/dimReduce.py:11-14), values drawn from a seeded generator and scaled so that every polynomial
approximation of the circuit is used inside its fitting interval (Config 1).

Matrix conventions follow what the circuit computes:
  *_weight_T.txt, ffn_W0_transposed_block_b.txt, pooler_dense_weight_T.txt : F with y = x . F      ("RE" products)
  selfAttn_WO_weight.txt, ffn_W2_block_b.txt, fcLinear_0_weight.txt        : F with y = F . x      ("CR" products)
"""
import math
import os

import numpy as np

D, K_PROJ, FFN, MAX_CLASSES = 128, 32, 512, 20
LAYER = "linformer_transformerLayers_transformer0_"


def _write(path, arr):
    a = np.atleast_2d(np.asarray(arr, np.float64))
    with open(path, "w") as f:
        for row in a:
            f.write(",".join("%.18e" % v for v in row) + "\n")


def make_model(seed=20261018, n_classes=20):
    """Random-init weights of the reference architecture (1 layer, d=128, k=32, FFN 512, <= 20 classes)."""
    rng = np.random.default_rng(seed)
    u = lambda shape, bound: rng.uniform(-bound, bound, shape)
    m = {
        "cls_token": np.clip(rng.normal(0, 0.25, D), -1, 1),
        "WQ_T": u((D, D), 3 / math.sqrt(D)), "bQ": u(D, 0.05),   # x3: gives the 32 attention scores a visible spread
        "WK_T": u((D, D), 3 / math.sqrt(D)), "bK": u(D, 0.05),
        "WV_T": u((D, D), 1 / math.sqrt(D)), "bV": u(D, 0.05),
        "WO": u((D, D), 1 / math.sqrt(D)), "bO": u(D, 0.05),
        "a1": rng.uniform(0.4, 0.6, D), "b1": u(D, 0.05), "c1": (1.0, 0.0, 0.0),   # keeps the bootstrap inputs inside [-1, 1]
        "W0_T": u((D, FFN), 1 / math.sqrt(D)), "b0": u(FFN, 0.05),
        "W2": u((D, FFN), 1 / math.sqrt(FFN)), "b2": u(D, 0.05),
        "a2": rng.uniform(0.8, 1.2, D), "b2n": u(D, 0.05), "c2": (1.0, 0.0, 0.0),
        "Wp_T": u((D, D), 1 / math.sqrt(D)), "bp": u(D, 0.05),
        "Wc": np.zeros((MAX_CLASSES, D)), "bc": np.zeros(MAX_CLASSES),
        "E": u((K_PROJ, 701), 1 / math.sqrt(701)), "Eb": u((K_PROJ, 1), 0.01),
        "F": u((K_PROJ, 701), 1 / math.sqrt(701)), "Fb": u((K_PROJ, 1), 0.01),
        "n_classes": n_classes,
    }
    m["Wc"][:n_classes] = u((n_classes, D), 1 / math.sqrt(D)) * 4.0   # rows n_classes..19 stay zero (main.cpp:121 reads 20 logits)
    m["bc"][:n_classes] = u(n_classes, 0.05)
    return m


def make_sample(model, tokens, seed):
    """One sample: `tokens` embeddings (S = tokens + 1 rows with CLS; the circuit needs 129 <= S <= 256) and the client-side
    Linformer projections X_E, X_F (src/python/dimReduce.py:153-160: E[:, :S] . X + E_b, over CLS + tokens)."""
    rng = np.random.default_rng(seed)
    x = np.clip(rng.normal(0, 0.25, (tokens, D)), -1, 1)
    rows = np.vstack([model["cls_token"][None, :], x])
    s = rows.shape[0]
    return {"tokens": x, "XE": model["E"][:, :s] @ rows + model["Eb"], "XF": model["F"][:, :s] @ rows + model["Fb"]}


def write_files(root, model, sample):
    """Lay the files out under `root` as the reference expects them relative to its working directory."""
    wd, ind, tok = (os.path.join(root, d) for d in ("weights-20NG", "input", "tokens"))
    for d in (wd, ind, tok, os.path.join(root, "keys"), os.path.join(root, "checkpoint")):
        os.makedirs(d, exist_ok=True)
    w = lambda name, arr: _write(os.path.join(wd, name), arr)
    w("cls_token.txt", model["cls_token"])
    for n in "QKV":
        w(LAYER + "selfAttn_W%s_weight_T.txt" % n, model["W%s_T" % n])
        w(LAYER + "selfAttn_W%s_bias.txt" % n, model["b" + n])
    for n in "EF":   # projection matrices, for the encrypted-projection variant (names of src/python/dimReduce.py:146-151)
        w(LAYER + "selfAttn_%s_weight.txt" % n, model[n])
        w(LAYER + "selfAttn_%s_bias.txt" % n, model[n + "b"].ravel())
    w(LAYER + "selfAttn_WO_weight.txt", model["WO"])
    w(LAYER + "selfAttn_WO_bias.txt", model["bO"])
    for idx, (a, b, c) in (("1", ("a1", "b1", "c1")), ("2", ("a2", "b2n", "c2"))):
        w(LAYER + "ffn_affine%s_a.txt" % idx, model[a])
        w(LAYER + "ffn_affine%s_b.txt" % idx, model[b])
        for k in range(3):
            w(LAYER + "ffn_affine%s_c%d.txt" % (idx, k), [model[c][k]])
    for b in range(4):
        w("ffn_W0_transposed_block_%d.txt" % b, model["W0_T"][:, 128 * b:128 * (b + 1)])
        w("ffn_W2_block_%d.txt" % b, model["W2"][:, 128 * b:128 * (b + 1)])
    w(LAYER + "ffn_Wffn_0_bias.txt", model["b0"])
    w(LAYER + "ffn_Wffn_2_bias.txt", model["b2"])
    w("pooler_dense_weight_T.txt", model["Wp_T"])
    w("pooler_dense_bias.txt", model["bp"])
    w("fcLinear_0_weight.txt", model["Wc"])
    w("fcLinear_0_bias.txt", np.concatenate([model["bc"], np.zeros(D - MAX_CLASSES)]))   # read as a 128-vector (F.cpp:651-672)
    write_sample_files(ind, tok, sample)
    return {"weights": wd, "input": ind, "tokens": tok}


def write_sample_files(input_dir, tokens_dir, sample):
    """The per-sample files only (../input/X{E,F}_i.txt M:161,166 and <input_folder>/input_i.txt M:172): a batch of samples shares
    one weights directory."""
    os.makedirs(input_dir, exist_ok=True); os.makedirs(tokens_dir, exist_ok=True)
    for i in range(K_PROJ):
        _write(os.path.join(input_dir, "XE_%d.txt" % i), sample["XE"][i])
        _write(os.path.join(input_dir, "XF_%d.txt" % i), sample["XF"][i])
    for f in os.listdir(tokens_dir):
        os.remove(os.path.join(tokens_dir, f))
    for i, row in enumerate(sample["tokens"]):
        _write(os.path.join(tokens_dir, "input_%d.txt" % i), row)
