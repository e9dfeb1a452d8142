"""torchrun entry: limb-sharded EvalRotate over NCCL vs the single-GPU rotation (bit-exact check + timing).
   python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/sharded_ks_check.py [logN] [l]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from fhe_linformer_b200 import Engine, sharded
logN = int(sys.argv[1]) if len(sys.argv) > 1 else 16
l = int(sys.argv[2]) if len(sys.argv) > 2 else 28
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
e = Engine(device=local, logN=logN)
sharded.register_signatures(e.lib)
dev = torch.device("cuda", local)
rng = np.random.default_rng(7)          # the same operands on every rank
ct = np.stack([np.stack([rng.integers(0, int(e.moduli[m]), e.N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
evk_h = rng.integers(0, 1 << 50, (e.dnum, 2, e.L + e.K, e.N), dtype=np.uint64)
evk = e.to_dev(evk_h); d_ct = e.to_dev(ct); t_ct = sharded.to_tensor(ct, dev)
g = e.galois(1)
want = e.rotate(d_ct, g, evk).download()
ks = sharded.ShardedKeySwitch(e, l, sharded.DistComm(), device=dev)
ks.rotate(t_ct, g, evk); got = ks.gather_result(); e.sync(); torch.cuda.synchronize()
ok = bool((got.cpu().numpy().view(np.uint64) == want).all())
ks_g = sharded.ShardedKeySwitch(e, l, sharded.DistComm(), device=dev, gather_digits=True)
ks_g.rotate(t_ct, g, evk); got_g = ks_g.gather_result(); e.sync(); torch.cuda.synchronize()
ok = ok and bool((got_g.cpu().numpy().view(np.uint64) == want).all())
stream = torch.cuda.ExternalStream(e.stream(), device=dev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    e.sync(); torch.cuda.synchronize(); dist.barrier()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    z.record(stream); e.sync(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(z) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
out = e.buf(d_ct.shape)
single = timed(lambda: e.rotate(d_ct, g, evk, out=out))
shard_ms = timed(lambda: ks.rotate(t_ct, g, evk))
shard_g_ms = timed(lambda: ks_g.rotate(t_ct, g, evk))
full_ms = timed(lambda: (ks.rotate(t_ct, g, evk), ks.gather_result()))
# what the exchanges alone cost: the same all-gathers without any kernel
comm = sharded.DistComm()
xchg_ms = timed(lambda: comm.all_gather(ks.states[0].pshare))
if rank == 0:
    print("limb-sharded EvalRotate N=2^%d l=%d over %d GPUs: bit-exact=%s  single-GPU %.1f us | sharded, result left limb-sharded: %.1f us (x%.2f, %.1f MB received per rank)"
          " | digits all-gathered too: %.1f us (x%.2f, %.1f MB) | + all-gather of the result: %.1f us (x%.2f) | the special-limb all-gather alone: %.1f us"
          % (logN, l, world, ok, single * 1e3, shard_ms * 1e3, single / shard_ms, ks.exchanged_bytes() / 1e6, shard_g_ms * 1e3, single / shard_g_ms,
             ks_g.exchanged_bytes() / 1e6, full_ms * 1e3, single / full_ms, xchg_ms * 1e3), flush=True)
dist.destroy_process_group()
