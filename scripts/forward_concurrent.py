"""Throughput experiment: T encrypted forwards in flight on ONE GPU (one FHEController + engine stream per host thread)."""
import os, sys, time, tempfile, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fhe_linformer_b200 import synth, host
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = 3
model = synth.make_model(n_classes=8)
ctl = []
for t in range(T):
    root = tempfile.mkdtemp(prefix="flb200_%d_" % t)
    sample = synth.make_sample(model, 128, seed=77 + t)
    dirs = synth.write_files(root, model, sample)
    fc = host.FHEController(root=root).generate()
    fc.forward(dirs)
    ctl.append((fc, dirs))
def work(fc, dirs, out):
    for _ in range(reps):
        out.append(fc.forward(dirs)[0])
outs = [[] for _ in range(T)]
th = [threading.Thread(target=work, args=(fc, dirs, outs[i])) for i, (fc, dirs) in enumerate(ctl)]
t0 = time.time()
for x in th: x.start()
for x in th: x.join()
dt = time.time() - t0
print("%d forwards in flight: %.3f samples/s (%.3f s per sample per stream), classes %s" % (T, T * reps / dt, dt / reps, [int(np.argmax(o[-1])) for o in outs]))
