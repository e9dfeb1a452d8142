"""BASELINE.json configs 1, 3, 4 (and the all-token variant) as encrypted forwards on one B200: seconds/sample, logit error against the
slot simulator and class agreement.  Synthetic weights / inputs of the named shape; reference CKKS parameters."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth, host
from oracle import linformer_sim as ls

CONFIGS = [  # name, classes, S, encrypted E/F projection, all-token attention
    ("config 1: R8-shaped sample (8 classes, S=200)", 8, 200, False, False),
    ("config 3: IMDB-shaped (2 classes, S=256), client-projected", 2, 256, False, False),
    ("config 3: IMDB-shaped (2 classes, S=256), E/F projection under encryption", 2, 256, True, False),
    ("config 4: BBC-shaped (5 classes, S=200)", 5, 200, False, False),
    ("config 4: 20NG-shaped (20 classes, S=256)", 20, 256, False, False),
    ("main_2.cpp variant: all-token attention (8 classes, S=129)", 8, 129, False, True),
]
root = tempfile.mkdtemp(prefix="flb200_")
fc = None
for name, ncls, S, encp, allt in CONFIGS:
    model = synth.make_model(n_classes=ncls)
    sample = synth.make_sample(model, S - 1, seed=3000 + S + ncls)
    dirs = synth.write_files(root, model, sample)
    if fc is None:
        fc = host.FHEController(root=root).generate()
    ref = ls.sim_forward(model, sample, all_tokens=allt) if allt else ls.sim_forward(model, sample)
    kw = dict(dead_work=True, encrypted_projection=encp, all_tokens=allt)
    fc.forward(dirs, **kw)                       # warm-up (mask / weight encodings, lazily generated keys)
    ts = []
    for _ in range(3):
        t = time.time(); logits, stages, toks = fc.forward(dirs, **kw); ts.append(time.time() - t)
    err = float(np.abs(logits[:ncls] - ref[:ncls]).max())
    print("%-78s %.3f s/sample (%s)  max logit err %.1e  class %d/%d" % (name, sorted(ts)[1], " ".join("%.3f" % x for x in ts), err,
          int(np.argmax(logits[:ncls])), int(np.argmax(ref[:ncls]))), flush=True)
