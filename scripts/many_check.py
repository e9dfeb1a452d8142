"""Several samples per packed forward (flh_forward_many): agreement with single-sample runs and samples/s as a function of the group size."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth, host
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
model = synth.make_model(n_classes=8)
root = tempfile.mkdtemp(prefix="flb200_many_")
samples = [synth.make_sample(model, S - 1, seed=100 + i) for i in range(8)]
dirs0 = synth.write_files(root, model, samples[0])
dirs = [dirs0]
for i, sm in enumerate(samples[1:]):
    d = {"weights": dirs0["weights"], "input": os.path.join(root, "input%d" % i), "tokens": os.path.join(root, "tokens%d" % i)}
    synth.write_sample_files(d["input"], d["tokens"], sm); dirs.append(d)
fc = host.FHEController(root=root).generate()
fc.set_option("packed_keys", 1)
single = [fc.forward(d, packed=True, dead_work=False)[0] for d in dirs]
for lean in (True, False):
    for M in (1, 2, 4, 8):
        fc.forward_many(dirs[:M], dead_work=not lean)
        ts = []
        for _ in range(3):
            t = time.time(); z, _ = fc.forward_many(dirs[:M], dead_work=not lean); ts.append(time.time() - t)
        dt = sorted(ts)[1]
        err = max(np.abs(z[m] - single[m]).max() for m in range(M))
        print("%s M=%d: %.3f s per call, %.2f samples/s, max |logit - single run| %.2e" % ("packed+lean" if lean else "packed", M, dt, M / dt, err), flush=True)
