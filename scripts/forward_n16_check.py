import os, sys, time, tempfile
sys.path.insert(0, "/root/repo")
import numpy as np
from fhe_linformer_b200 import synth, host
from oracle import linformer_sim as ls
model = synth.make_model(n_classes=8); sample = synth.make_sample(model, 128, seed=20261018 + 1)
root = tempfile.mkdtemp(prefix="flb200_"); dirs = synth.write_files(root, model, sample)
t = time.time(); fc = host.FHEController(root=root).generate(log_ring=16); print("keys %.1fs N=%d" % (time.time() - t, fc.ckks.N))
v = np.random.default_rng(0).uniform(-1, 1, 16384)
ct = fc.ckks.encrypt(v, level=24, slots=16384)
b = fc.ckks.bootstrap(ct); print("sparse bootstrap err", np.abs(fc.ckks.decrypt(b) - v).max(), b.level)
for i in range(2):
    t = time.time(); logits, stages, S = fc.forward(dirs); print("forward N=2^16: %.2fs" % (time.time() - t), stages)
ref = ls.sim_forward(model, sample); print("max logit err", np.abs(logits - ref).max(), int(np.argmax(logits)), int(np.argmax(ref)))
