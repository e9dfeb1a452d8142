"""Writes the files /root/reference/src/python/compute_simple.py reads from this repo's synthetic model (untransposed weights,
a zero positional table, the 32 x 701 E / F matrices, one value per line for the token embeddings), runs the reference script
on them (run_reference_python.py) and stores inputs' seeds + the script's outputs as tests/golden/ref_python_model.npz.

    python tests/golden/make_ref_python_golden.py            # needs /root/reference; the fixture travels, the reference does not
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fhe_linformer_b200 import synth  # noqa: E402

REF_SCRIPT = "/root/reference/src/python/compute_simple.py"
LAYER = synth.LAYER
MODEL_SEED, SAMPLE_SEED, TOKENS, CLASSES = 20261018, 7, 140, 20


def write_reference_files(root, model, sample):
    """File names and shapes of compute_simple.py:129-225."""
    wd, tok = os.path.join(root, "weights"), os.path.join(root, "tokens")
    os.makedirs(wd, exist_ok=True); os.makedirs(tok, exist_ok=True)
    w = lambda name, arr: synth._write(os.path.join(wd, name), arr)
    w("posEmb.txt", np.zeros((700, synth.D)))                       # :129 -- zero: the C++ circuit adds none to the token rows
    w("cls_token.txt", model["cls_token"])
    for n in "EF":
        w(LAYER + "selfAttn_%s_weight.txt" % n, model[n])           # 32 x 701
        w(LAYER + "selfAttn_%s_bias.txt" % n, model[n + "b"].ravel())
    for n in "QKV":
        w(LAYER + "selfAttn_W%s_weight.txt" % n, model["W%s_T" % n].T)   # untransposed: the script computes x @ W.T
        w(LAYER + "selfAttn_W%s_bias.txt" % n, model["b" + n])
    w(LAYER + "selfAttn_WO_weight.txt", model["WO"])
    w(LAYER + "selfAttn_WO_bias.txt", model["bO"])
    for idx, (a, b, c) in (("1", ("a1", "b1", "c1")), ("2", ("a2", "b2n", "c2"))):
        w(LAYER + "ffn_affine%s_a.txt" % idx, model[a])
        w(LAYER + "ffn_affine%s_b.txt" % idx, model[b])
        for k in range(3):
            w(LAYER + "ffn_affine%s_c%d.txt" % (idx, k), [model[c][k]])
    w(LAYER + "ffn_Wffn_0_weight.txt", model["W0_T"].T)             # 512 x 128
    w(LAYER + "ffn_Wffn_0_bias.txt", model["b0"])
    w(LAYER + "ffn_Wffn_2_weight.txt", model["W2"])                 # 128 x 512
    w(LAYER + "ffn_Wffn_2_bias.txt", model["b2"])
    w("pooler_dense_weight.txt", model["Wp_T"].T)
    w("pooler_dense_bias.txt", model["bp"])
    w("fcLinear_0_weight.txt", model["Wc"])
    w("fcLinear_0_bias.txt", model["bc"])
    for i, row in enumerate(sample["tokens"]):                       # np.loadtxt without a delimiter: one value per line
        with open(os.path.join(tok, "input_%d.txt" % i), "w") as f:
            f.write("\n".join("%.18e" % v for v in row) + "\n")
    return tok, wd


def run_reference(model, sample, workdir):
    tok, wd = write_reference_files(workdir, model, sample)
    out = os.path.join(workdir, "ref_out.npz")
    subprocess.run([sys.executable, os.path.join(HERE, "run_reference_python.py"), REF_SCRIPT, tok, wd, out], check=True, stdout=subprocess.DEVNULL,
                   cwd=os.path.dirname(REF_SCRIPT))
    return dict(np.load(out))


def main():
    model = synth.make_model(MODEL_SEED, CLASSES)
    sample = synth.make_sample(model, TOKENS, SAMPLE_SEED)
    with tempfile.TemporaryDirectory() as d:
        ref = run_reference(model, sample, d)
    keep = {k: ref[k] for k in ("K", "Q", "exp_approx", "attn_out", "x_norm0", "x_norm1", "y_logit", "y_pred")}
    keep["Q"] = keep["Q"][:1]            # only Q[0] is read downstream (compute_simple.py:162)
    keep["x_norm0"], keep["x_norm1"] = keep["x_norm0"][:2], keep["x_norm1"][:2]
    np.savez_compressed(os.path.join(HERE, "ref_python_model.npz"), model_seed=MODEL_SEED, sample_seed=SAMPLE_SEED, tokens=TOKENS, classes=CLASSES,
                        **{k: np.asarray(v, np.float64) for k, v in keep.items()})
    print("wrote ref_python_model.npz: logits", np.round(keep["y_logit"][0][:6], 4), "pred", int(keep["y_pred"]))


if __name__ == "__main__":
    main()
