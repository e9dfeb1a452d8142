"""Writes tests/golden/linformer_sim_s129.npz: logits and a few checkpoint digests of the slot simulator for the seeded
synthetic model (committed fixture; regenerate with `python scripts/make_golden_linformer.py`).  The reference itself holds
no golden vectors, so this pins OUR restatement against accidental change, nothing more."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth
from oracle import linformer_sim as ls

model = synth.make_model(n_classes=8)
sample = synth.make_sample(model, 128, seed=20261018 + 1)
cp = {}
logits = ls.sim_forward(model, sample, cp)
out = {"logits": logits}
for k in ("scores_exp", "attention_cls", "affine1_0", "container0_gelu", "encoder_out", "pooler_out"):
    out["cp_" + k] = cp[k][:: 97].copy()   # strided digest keeps the fixture small
np.savez_compressed(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "linformer_sim_s129.npz"), **out)
print({k: v.shape for k, v in out.items()})
