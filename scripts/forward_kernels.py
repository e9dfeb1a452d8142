"""Kernel-level breakdown driver: generate keys, one warm-up forward, then one forward bracketed by two marker launches
(fl_sync + a 1-limb NTT of a scratch polynomial) so an ncu launch list can be cut at the markers."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth, host
S = int(sys.argv[1]) if len(sys.argv) > 1 else 129
model = synth.make_model(n_classes=8); sample = synth.make_sample(model, S - 1, seed=5)
root = tempfile.mkdtemp(prefix="flb200_"); dirs = synth.write_files(root, model, sample)
fc = host.FHEController(root=root).generate()
fc.forward(dirs)
c = fc.ckks
marker = lambda: c.decrypt(c.encrypt(np.ones(4), level=27, slots=4))      # level 27 = one limb: a recognisable tiny launch group
marker()
logits, stages, toks = fc.forward(dirs)
marker()
print("done", toks)
