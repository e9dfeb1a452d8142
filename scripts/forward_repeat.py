import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhe_linformer_b200 import synth, host
model = synth.make_model(n_classes=8); sample = synth.make_sample(model, 128, seed=5)
root = tempfile.mkdtemp(prefix="flb200_"); dirs = synth.write_files(root, model, sample)
fc = host.FHEController(root=root).generate()
for i in range(8):
    if i == 4: fc.ckks.ledger(True)
    t = time.time(); logits, stages, toks = fc.forward(dirs); dt = time.time() - t
    import ctypes as C
    info = (C.c_int * 8)(); fc.ckks.lib.fl_ctx_info(fc.ckks.h, info)
    print("run %d%s: %.3f s %s | pool allocs %d trims %d cached %d MiB" % (i, " (ledger on)" if i >= 4 else "", dt, {k: round(v, 3) for k, v in stages.items()}, info[5], info[6], info[7]), flush=True)
