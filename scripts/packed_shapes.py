"""Packed / packed + lean forwards at the other BASELINE shapes (S = 129, 256; 2, 5, 20 classes) against the slot simulator."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth, host
from oracle import linformer_sim as ls
fc = None
for classes, S, encp in ((2, 256, False), (2, 256, True), (5, 200, False), (20, 256, False), (8, 129, False)):
    model = synth.make_model(n_classes=classes); sample = synth.make_sample(model, S - 1, seed=9 + S + classes)
    root = tempfile.mkdtemp(prefix="flb200_shapes_"); dirs = synth.write_files(root, model, sample)
    if fc is None:
        fc = host.FHEController(root=root).generate(); fc.set_option("packed_keys", 1)
    ref = ls.sim_forward(model, sample)
    for lean in (False, True):
        fc.forward(dirs, packed=True, dead_work=not lean, encrypted_projection=encp)
        t = time.time(); z, _, toks = fc.forward(dirs, packed=True, dead_work=not lean, encrypted_projection=encp); dt = time.time() - t
        print("classes %2d S %3d enc-proj %d %-11s: %.3f s, max |logit - simulator| %.2e, class %d / %d" % (classes, toks, encp, "packed+lean" if lean else "packed", dt,
              np.abs(z - ref).max(), int(np.argmax(z[:classes])), int(np.argmax(ref[:classes]))), flush=True)
try:
    fc.forward(dirs, packed=True, all_tokens=True)
except RuntimeError as e:
    print("packed + all-token:", e)
