"""Profiling driver: ONE encrypted forward (S = argv[1], mode argv[2] in faithful|packed|packed_lean) between cudaProfilerStart/Stop,
for `ncu --profile-from-start off --metrics gpu__time_duration.sum` launch lists."""
import ctypes, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhe_linformer_b200 import synth, host
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
mode = sys.argv[2] if len(sys.argv) > 2 else "packed_lean"
model = synth.make_model(n_classes=8); sample = synth.make_sample(model, S - 1, seed=5)
root = tempfile.mkdtemp(prefix="flb200_"); dirs = synth.write_files(root, model, sample)
fc = host.FHEController(root=root).generate()
kw = dict(packed=mode.startswith("packed"), dead_work=not mode.endswith("lean"))
if kw["packed"]: fc.set_option("packed_keys", 1)
fc.forward(dirs, **kw); fc.forward(dirs, **kw)
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
fc.forward(dirs, **kw); fc.ckks.sync()
rt.cudaProfilerStop()
print("done")
