"""Encrypted forward on the GPU vs the slot simulator, checkpoint by checkpoint (developer script; the test is tests/test_gpu_forward.py)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth, host
from oracle import linformer_sim as ls

S = int(sys.argv[1]) if len(sys.argv) > 1 else 129
dead = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model = synth.make_model(n_classes=8)
sample = synth.make_sample(model, S - 1, seed=20261018 + 1)
root = tempfile.mkdtemp(prefix="flb200_")
dirs = synth.write_files(root, model, sample)
t = time.time()
fc = host.FHEController(root=root).generate()
print("context + keys: %.2fs, rot keys %d" % (time.time() - t, fc.ckks.num_rot_keys()), flush=True)
fc.ckks.ledger(True)
cp = {}
t = time.time()
logits, stages, toks = fc.forward(dirs, dead_work=bool(dead), checkpoints=cp)
dt = time.time() - t
print("forward S=%d: %.2fs" % (toks, dt), stages)
ref_cp = {}
ref = ls.sim_forward(model, sample, ref_cp)
for name, (v, level) in cp.items():
    r = ref_cp.get(name)
    if r is None: continue
    print("%-22s lvl %2d  max|ref| %.4f  max err %.3e" % (name, level, np.abs(r).max(), np.abs(v - r).max()))
print("logits  ", np.round(logits[:8], 5))
print("expected", np.round(ref[:8], 5))
print("max logit err %.3e  argmax %d vs %d" % (np.abs(logits - ref).max(), int(np.argmax(logits)), int(np.argmax(ref))))
led = fc.ckks.ledger_dump()
rot = sum(n for k, (n, b) in led.items() if k.startswith("rotate@")); tot_b = sum(b for n, b in led.values())
print("ledger: %d rotations, %d entries, %.1f GB algorithmic -> %.1f GB/s" % (rot, len(led), tot_b / 1e9, tot_b / 1e9 / dt))
