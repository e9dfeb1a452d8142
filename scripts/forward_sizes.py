"""Forward seconds/sample at S = 129, 200, 256 (faithful and lean), reference parameters (developer script)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from fhe_linformer_b200 import synth, host
from oracle import linformer_sim as ls
model = synth.make_model(n_classes=8)
root = tempfile.mkdtemp(prefix="flb200_")
fc = None
for S in (129, 200, 256):
    sample = synth.make_sample(model, S - 1, seed=1000 + S)
    dirs = synth.write_files(root, model, sample)
    if fc is None: fc = host.FHEController(root=root).generate()
    ref = ls.sim_forward(model, sample)
    for dead in (True, False):
        fc.forward(dirs, dead_work=dead)
        ts = []
        for _ in range(3):
            t = time.time(); logits, stages, toks = fc.forward(dirs, dead_work=dead); ts.append(time.time() - t)
        info = (C.c_int * 8)(); fc.ckks.lib.fl_ctx_info(fc.ckks.h, info)
        print("S=%d %-8s median %.3f s (%s)  max logit err %.1e  class %d/%d  cached %.1f GB" % (toks, "faithful" if dead else "lean", sorted(ts)[1], " ".join("%.3f" % x for x in ts), np.abs(logits - ref).max(), int(np.argmax(logits)), int(np.argmax(ref)), info[7] / 1024), flush=True)
