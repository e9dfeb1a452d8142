"""Where the encrypted forward spends GPU and host time, per C-ABI entry point (developer script)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth, host
S = int(sys.argv[1]) if len(sys.argv) > 1 else 129
packed = len(sys.argv) > 2 and sys.argv[2].startswith("packed")
lean = len(sys.argv) > 2 and sys.argv[2].endswith("lean")
model = synth.make_model(n_classes=8); sample = synth.make_sample(model, S - 1, seed=5)
root = tempfile.mkdtemp(prefix="flb200_"); dirs = synth.write_files(root, model, sample)
fc = host.FHEController(root=root).generate()
if packed: fc.set_option("packed_keys", 1)
fc.forward(dirs, packed=packed, dead_work=not lean); fc.forward(dirs, packed=packed, dead_work=not lean)                    # warm-up
fc.ckks.prof(True); fc.ckks.prof_dump()
t = time.time(); logits, stages, toks = fc.forward(dirs, packed=packed, dead_work=not lean); dt = time.time() - t
p = fc.ckks.prof_dump()
print("forward S=%d: %.3f s  stages %s" % (toks, dt, {k: round(v, 3) for k, v in stages.items()}))
tot_g = sum(v[1] for v in p.values()); tot_h = sum(v[2] for v in p.values())
print("%-22s %7s %10s %10s" % ("entry point", "calls", "gpu ms", "host ms"))
for k, (n, g, h) in sorted(p.items(), key=lambda kv: -kv[1][1]):
    print("%-22s %7d %10.1f %10.1f" % (k, n, g, h))
print("total gpu %.1f ms, host-in-call %.1f ms, wall %.1f ms" % (tot_g, tot_h, dt * 1e3))
