import os, sys, time, tempfile
sys.path.insert(0, "/root/repo")
import numpy as np
from fhe_linformer_b200 import synth, host
from oracle import linformer_sim as ls
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
root = tempfile.mkdtemp(prefix="flb200_")
model = synth.make_model(n_classes=8)
sample = synth.make_sample(model, S - 1, seed=77)
dirs = synth.write_files(root, model, sample)
fc = host.FHEController(root=root).generate()
t = time.time(); fc.set_option("packed_keys", 1); print("packed keys %.2f s, %.1f GB of rotation keys" % (time.time() - t, fc.rotation_key_bytes() / 1e9), flush=True)
ref = ls.sim_forward(model, sample)
cps = {}
lg, st, toks = fc.forward(dirs, packed=True, checkpoints=cps)
print("packed first run: max logit err %.2e class %d/%d" % (np.abs(lg[:8] - ref[:8]).max(), np.argmax(lg[:8]), np.argmax(ref[:8])), flush=True)
for k, (v, lvl) in cps.items(): print("  cp %-24s level %2d  max|v| %.3f" % (k, lvl, np.abs(v).max()))
fc.ckks.ledger(True); fc.ckks.ledger_reset()
ts = []
for _ in range(3):
    t = time.time(); lg, st, _ = fc.forward(dirs, packed=True); ts.append(time.time() - t)
led = fc.ckks.ledger_dump(); fc.ckks.ledger(False)
print("packed: %.3f s/sample (%s) err %.2e; rotations %d" % (sorted(ts)[1], " ".join("%.3f" % x for x in ts), np.abs(lg[:8] - ref[:8]).max(), sum(n for k, (n, _) in led.items() if k.startswith("rotate@")) // 3))
print("   stages", {k: round(v, 3) for k, v in st.items()})
cps = {}
lgl, stl, _ = fc.forward(dirs, packed=True, dead_work=False, checkpoints=cps)
print("packed+lean first: err %.2e class %d" % (np.abs(lgl[:8] - ref[:8]).max(), np.argmax(lgl[:8])))
for k, (v, lvl) in cps.items():
    if k.startswith(("packed", "encoder", "pooler")): print("  cp %-24s level %2d  max|v| %.3f" % (k, lvl, np.abs(v).max()))
ts = []
for _ in range(3):
    t = time.time(); lgl, stl, _ = fc.forward(dirs, packed=True, dead_work=False); ts.append(time.time() - t)
print("packed+lean: %.3f s/sample (%s) err %.2e" % (sorted(ts)[1], " ".join("%.3f" % x for x in ts), np.abs(lgl[:8] - ref[:8]).max()))
print("   stages", {k: round(v, 3) for k, v in stl.items()})
fc.forward(dirs, dead_work=True)
ts = []
for _ in range(2):
    t = time.time(); lg2, st2, _ = fc.forward(dirs, dead_work=True); ts.append(time.time() - t)
print("faithful: %.3f s/sample err %.2e;  packed vs faithful logits %.2e" % (min(ts), np.abs(lg2[:8] - ref[:8]).max(), np.abs(lg2 - lg).max()))
print("   stages", {k: round(v, 3) for k, v in st2.items()})
