#!/usr/bin/env python
"""bench.py -- BASELINE.json config[1]: CKKS primitive bench, EvalRotate at N=2^16 over the full RNS chain
(28 Q limbs + 7 P limbs, dnum=4), rotations/s; one step = one pass of EvalRotate over a batch of B
independent ciphertexts (ciphertext-parallel across ranks, no data-path collective).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EvalRotate throughput, N=2^16, full chain (L=28 Q limbs + 7 P limbs, dnum=4)"
UNIT = "rotations/s"


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout: everything else this process (or NCCL, or the controller's progress messages)
    writes to fd 1 is sent to stderr; emit() writes the line to the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--logN", type=int, default=16)
    ap.add_argument("--limbs", type=int, default=28, help="active Q limbs of the operands (28 = full chain)")
    ap.add_argument("--batch", type=int, default=32, help="ciphertexts per step per GPU")
    ap.add_argument("--group", type=int, default=32, help="ciphertexts sharing one key per batched call (the reference's row loops rotate S = 200 rows by one index)")
    ap.add_argument("--cpu-sample", type=int, default=20, help="rotations timed on the host for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-forward", action="store_true", help="skip the encrypted Linformer forward (extra `forward` key)")
    ap.add_argument("--forward-in-flight", type=int, default=2, help="forwards in flight per GPU for the extra throughput-mode figure (1 = skip)")
    ap.add_argument("--forward-rows", type=int, default=200, help="S = rows of the forward sample (129..256); 200 = SURVEY.md Config 1")
    ap.add_argument("--samples", type=int, default=0, help="samples of the sample-parallel batch over ALL GPUs (BASELINE config 5 names 256); 0 = 8 per GPU")
    ap.add_argument("--samples-per-call", type=int, default=4, help="samples per packed forward in the batched config-5 figure (forward.batch.packed_many)")
    ap.add_argument("--no-forward-configs", action="store_true", help="skip the BASELINE config 3 / 4 shapes (extra `forward.configs` block)")
    ap.add_argument("--no-forward-n16", action="store_true", help="skip the extra forward at the reference's commented-out ring N=2^16")
    ap.add_argument("--forward-logn", type=int, default=15, help="ring of the forward: 15 = the reference's parameters, 16 = its commented-out variant (sparse packing)")
    return ap.parse_args()


def host_threads():
    """All host cores this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which round 1's
    reference arm silently obeyed at N >= 2: the oracle's thread count is now set explicitly (orc_set_threads)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(logN, l, K, dnum, B, G, world, N):
    """`config` of the JSON line: the same object for both arms, so the driver's same_config check compares like with like."""
    return {"workload": f"EvalRotate N=2^{logN} l={l} K={K} dnum={dnum} (BASELINE.json configs[1])", "batch_per_gpu": B,
            "parallelism": f"ciphertext-parallel x{world}", "ciphertexts_per_launch": G,
            "l2": f"inputs larger than L2 ({B * 2 * 2 * l * N * 8 / 1e6:.0f} MB touched per step)"}


def algorithmic_bytes_rotate(N, l, K, alpha):
    beta = (l + alpha - 1) // alpha
    return (4 * l + 2 * beta * (l + K)) * 8 * N


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_load = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs a few hundred ms to produce its first line: wait for it so a short timed region is still covered."""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        """Clocks over the samples taken inside [t_begin, t_end] (the timed region); when the region is shorter than the
        sampling period, over the GPU-busy window that started with the warm-up steps of the same workload (`window`)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        ok = [(t, r) for t, r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        inside = [(t, r) for t, r in ok if t_begin is not None and t_begin - 0.05 <= t <= t_end + 0.12]
        window = "timed region"
        if not inside:
            inside, window = [(t, r) for t, r in ok if self.t_load is None or t >= self.t_load], "warm-up + timed region (timed region shorter than the sampling period)"
        sm = sorted(float(r[1]) for _, r in inside)
        mx = [float(r[2]) for _, r in inside if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, r in inside:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def cpu_rotations(o, l, count, seed=0, warm=True):
    """Oracle EvalRotate on the host (all OpenMP threads): returns seconds per rotation."""
    rng = np.random.default_rng(seed)
    ct = np.stack([np.stack([rng.integers(0, int(o.moduli[m]), o.N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
    evk = rng.integers(0, 1 << 50, (o.dnum, 2, o.L + o.K, o.N), dtype=np.uint64)
    g = o.galois(1)
    if warm:
        o.rotate(ct, g, evk)   # warm-up (page faults, twiddles into cache)
    t0 = time.perf_counter()
    for _ in range(count):
        o.rotate(ct, g, evk)
    return (time.perf_counter() - t0) / count


def run_reference(a):
    """--impl reference: the reference's CPU path.  OpenFHE cannot be built here, so this arm
    times the oracle port (oracle/ckks_oracle.c, OpenMP over all host cores) on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_threads())   # before libgomp initialises; orc_set_threads below covers the other order
    from oracle.oracle import Oracle, lib
    lib().orc_set_threads(host_threads())
    o = Oracle(logN=a.logN, L=28, dnum=4)
    cores = int(lib().orc_num_threads())
    per_step = max(1, min(a.batch, 4))                   # bounded sample of the step's batch
    for _ in range(a.warmup):
        cpu_rotations(o, a.limbs, 1, warm=False)
    secs = [cpu_rotations(o, a.limbs, per_step, seed=100 + s, warm=False) for s in range(a.steps)]
    per_rot = float(np.mean(secs))
    val = 1.0 / per_rot
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": per_rot * per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(a.logN, a.limbs, o.K, o.dnum, a.batch, max(1, min(a.group, a.batch)), a.gpus, o.N),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} of the step's {a.batch} rotations per step x {a.steps} steps, oracle/ckks_oracle.c with OpenMP ({cores} threads); "
                                   "OpenFHE-equivalent CPU restatement, not OpenFHE"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_b200(a):
    import torch
    import torch.distributed as dist
    from fhe_linformer_b200 import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    e = Engine(device=local, logN=a.logN)
    N, l, B = e.N, a.limbs, a.batch
    rng = np.random.default_rng(1234 + rank)
    q = e.moduli

    def rand_limbs(midx):
        return np.stack([rng.integers(0, int(q[m]), N, dtype=np.uint64) for m in midx])

    # evaluation keys: uniformly random residues (timing does not need a valid key); 2 keys alternate
    nkeys = 2
    evks = [e.to_dev(np.stack([rand_limbs(range(e.L + e.K)) for _ in range(e.dnum * 2)]).reshape(e.dnum, 2, e.L + e.K, N)) for _ in range(nkeys)]
    g = e.galois(1)
    base = np.stack([rand_limbs(range(l)), rand_limbs(range(l))])
    # the step's batch: B independent ciphertexts stored back to back, rotated in groups that share a key (the
    # ciphertext-parallel form of the reference's row loops); G ciphertexts per batched call
    G = max(1, min(a.group, B))
    host_batch = np.stack([np.roll(base, i + 1, axis=2) for i in range(B)])      # distinct contents, cheap to generate
    ins = [e.to_dev(host_batch[i:i + G]) for i in range(0, B, G)]
    outs = [e.buf(x.shape) for x in ins]
    stream = torch.cuda.ExternalStream(e.stream(), device=torch.device("cuda", local))

    def step():
        for i, x in enumerate(ins):
            e.rotate_batch(x, g, evks[i % nkeys], out=outs[i])

    def barrier():
        e.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first()
        sampler.t_load = time.time()
    for _ in range(a.warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    ev0.record(stream)
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    e.sync()
    t_end = time.time()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    from fhe_linformer_b200 import shard
    units, worst_s, value = shard.combine(B * a.steps, ms * 1e-3, device="cuda")     # SUM of rotations / MAX device time
    ms = worst_s * 1e3

    # ---- dominant kernel family: the NTT pass pair (column + chunk), timed alone over the same buffers ----
    midx2 = np.concatenate([np.arange(l), np.arange(l)]).astype(np.int32)
    # the forward transform of both polynomials of G ciphertexts per launch pair, B ciphertexts in total (> L2)
    scratch = [e.to_dev(host_batch[i:i + G].reshape(-1, 2 * l, N)) for i in range(0, B, G)]
    for sbuf in scratch:
        e.ntt_batch(sbuf, midx2)
    e.sync()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    n0.record(stream)
    for _ in range(reps):
        for sbuf in scratch:
            e.ntt_batch(sbuf, midx2)
    n1.record(stream)
    e.sync()
    ntt_ms = n0.elapsed_time(n1) / (reps * B)
    ntt_bytes = 16 * N * 2 * l
    # the other primitive rates of BASELINE.json's metric (ii): INTT and EvalMult + relinearisation, same buffers / chain
    n0.record(stream)
    for _ in range(reps):
        for sbuf in scratch:
            e.ntt_batch(sbuf, midx2, inverse=True)
    n1.record(stream)
    e.sync()
    intt_ms = n0.elapsed_time(n1) / (reps * B)
    ma, mb = e.to_dev(host_batch[0]), e.to_dev(host_batch[1 % B])
    mo = e.buf(ma.shape)
    e.mul_relin(ma, mb, evks[0], out=mo)
    e.sync()
    n0.record(stream)
    for _ in range(4 * reps):
        e.mul_relin(ma, mb, evks[0], out=mo)
    n1.record(stream)
    e.sync()
    mul_ms = n0.elapsed_time(n1) / (4 * reps)
    beta_l = (l + e.alpha - 1) // e.alpha
    mul_bytes = (6 * l + 2 * beta_l * (l + e.K)) * 8 * N
    for x in (ma, mb, mo):
        x.free()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ntt_gbs = ntt_bytes / (ntt_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch of the same kernel pair from the committed ncu capture (scaled to this limb count)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["ntt_pass_pair_2x28_limbs_per_ciphertext"]
        if a.logN == 16:
            traffic = tj["dram_bytes_per_launch"] * l / 28.0 * G
    except Exception:
        pass
    rot_bytes = algorithmic_bytes_rotate(N, l, e.K, e.alpha)
    rot_gbs = rot_bytes * value / world / 1e9
    for s in scratch:
        s.free()

    # ---- N > 1 only: ONE EvalRotate split across the ranks by limb (the optional limb-sharded key switch, all-gathers over NVLink) ----
    sharded_blk = None
    if world > 1:
        from fhe_linformer_b200 import sharded
        sharded.register_signatures(e.lib)
        dev = torch.device("cuda", local)
        rs = np.random.default_rng(7)                          # the same operands on every rank
        ct1 = np.stack([np.stack([rs.integers(0, int(q[m]), N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
        evk1 = e.to_dev(rs.integers(0, 1 << 50, (e.dnum, 2, e.L + e.K, N), dtype=np.uint64))
        d1, t1 = e.to_dev(ct1), sharded.to_tensor(ct1, dev)
        o1 = e.buf(d1.shape)
        want = e.rotate(d1, g, evk1).download()

        def timed_ms(fn, reps=20):
            for _ in range(3):
                fn()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(reps):
                fn()
            a1.record(stream); e.sync(); torch.cuda.synchronize()
            tt = torch.tensor([a0.elapsed_time(a1) / reps], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        rows = {}
        ok_all = True
        for name, gd in (("digits recomputed by every rank", False), ("digits all-gathered", True)):
            ks = sharded.ShardedKeySwitch(e, l, sharded.DistComm(), device=dev, gather_digits=gd)
            ks.rotate(t1, g, evk1)
            got1 = ks.gather_result(); e.sync(); torch.cuda.synchronize()
            ok_all = ok_all and bool((got1.cpu().numpy().view(np.uint64) == want).all())
            rows[name] = {"us": timed_ms(lambda: ks.rotate(t1, g, evk1)) * 1e3, "MB_received_per_rank": ks.exchanged_bytes() / 1e6}
            if not gd:
                rows["... plus an all-gather of the result limbs"] = {"us": timed_ms(lambda: (ks.rotate(t1, g, evk1), ks.gather_result())) * 1e3}
                comm = sharded.DistComm()
                xchg_us = timed_ms(lambda: comm.all_gather(ks.states[0].pshare)) * 1e3
        single_us = timed_ms(lambda: e.rotate(d1, g, evk1, out=o1)) * 1e3
        okt = torch.tensor([1.0 if ok_all else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        best = min(v["us"] for k, v in rows.items() if not k.startswith("..."))
        sharded_blk = {"ranks": world, "ring": f"N=2^{a.logN}, l={l}", "bit_exact_vs_single_gpu": bool(okt.item() > 0.5), "single_gpu_us": single_us,
                       "sharded": rows, "latency_ratio_single_over_sharded": single_us / best, "special_limb_all_gather_alone_us": xchg_us,
                       "note": "one EvalRotate (no batching), limbs of Q_l u P split across the ranks, result left limb-sharded; CUDA events on the engine "
                               "stream, max over ranks; fhe_linformer_b200/sharded.py"}
        for x in (d1, o1, evk1):
            x.free()

    # ---- e2e: same metric through the host-buffer C-ABI call (H2D + kernels + D2H inside the timed region) ----
    ngroups = len(ins)
    pin_in = [torch.empty(tuple(x.shape), dtype=torch.int64).pin_memory() for x in ins]
    pin_out = [torch.empty(tuple(x.shape), dtype=torch.int64).pin_memory() for x in ins]
    h_in = [p.numpy().view(np.uint64) for p in pin_in]
    h_out = [p.numpy().view(np.uint64) for p in pin_out]
    for i in range(ngroups):
        h_in[i][...] = host_batch[i * G:(i + 1) * G]
    e2e_steps = max(2, min(a.steps, 5))
    e.host_rotate_batch(h_in[0], g, evks[0], out=h_out[0])
    e.host_rotate_batch(h_in[0], g, evks[0], out=h_out[0], wait=False)    # warm-up of the overlapped form (its batch shapes, both staging slots)
    e.sync()
    e2e_runs = []
    for _rep in range(3):     # the PCIe path is shared with the other tenants of the host: three timed regions, the median is reported
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            for i in range(ngroups):
                e.host_rotate_batch(h_in[i], g, evks[i % nkeys], out=h_out[i], wait=False)   # calls overlap; one wait at the end of the timed region
        e.sync()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_runs.append(world * B * e2e_steps / float(te.item()))
    e2e_val = sorted(e2e_runs)[1]
    # what the link gives at that moment: the same pinned buffers copied both ways at once, no kernels
    cp0, cp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    d_t = torch.empty(tuple(ins[0].shape), dtype=torch.int64, device="cuda")
    d_t2 = torch.empty_like(d_t)
    torch.cuda.synchronize()
    tp = time.perf_counter()
    for _ in range(4):
        with torch.cuda.stream(s_up):
            d_t.copy_(pin_in[0], non_blocking=True)
        with torch.cuda.stream(s_dn):
            pin_out[1 % ngroups].copy_(d_t2, non_blocking=True)
    torch.cuda.synchronize()
    pcie_gbs = 4 * d_t.numel() * 8 / (time.perf_counter() - tp) / 1e9
    del d_t, d_t2
    e.host_rotate_batch(h_in[0], g, evks[0], out=h_out[0])               # h_out[0] again holds the result checked below
    ok = bool((h_out[0] == outs[0].download()).all())      # e2e result equals the device-resident result

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle.oracle import Oracle, lib
        lib().orc_set_threads(host_threads())
        o = Oracle(logN=a.logN, L=28, dnum=4)
        per = cpu_rotations(o, l, a.cpu_sample)
        cores = int(lib().orc_num_threads())
        cpu = {"value": 1.0 / per, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{a.cpu_sample} EvalRotate at N=2^{a.logN}, l={l}, oracle/ckks_oracle.c OpenMP {cores} threads "
                         "(OpenFHE-equivalent CPU restatement, not OpenFHE)"}

    fwd = None
    if not a.no_forward:
        fwd = run_forward(a, local, rank, world, torch, dist)
        if fwd and cpu is not None:
            fwd["cpu_estimate"] = forward_cpu_estimate(a, fwd.pop("_ledger"), e.K, e.alpha)
        if fwd:
            fwd.pop("_ledger", None)

    if rank == 0:
        launches_per_group = 11         # kernels per batched call: 4 transform pass pairs (8; the last one carries the ModDown finish) + ModUp conversion, key inner product, ModDown conversion
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": workload_config(a.logN, l, e.K, e.dnum, B, G, world, N),
            "roofline": {"bound": "hbm", "kernel": "ntt pass pair (ntt_column_kernel + ntt_chunk_kernel)", "achieved": ntt_gbs, "peak": peak,
                         "unit": "GB/s", "frac": ntt_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ntt_bytes * G, "avg_launch_ms": ntt_ms * G,
                         "units_per_launch": f"{G} ciphertexts x {2 * l} limbs (16 N bytes per limb)"},
            "rotate_roofline": {"algorithmic_bytes_per_rotation": rot_bytes, "achieved": rot_gbs, "unit": "GB/s", "frac": rot_gbs / peak},
            "primitives": {"ring": f"N=2^{a.logN}, l={l}", "ntt_polys_per_s": 2.0 / (ntt_ms * 1e-3), "intt_polys_per_s": 2.0 / (intt_ms * 1e-3),
                           "ntt_limbs_per_s": 2.0 * l / (ntt_ms * 1e-3), "intt_GBps": ntt_bytes / (intt_ms * 1e-3) / 1e9,
                           "mul_relin_per_s": 1e3 / mul_ms, "mul_relin_GBps": mul_bytes / (mul_ms * 1e-3) / 1e9,
                           "mul_relin_roofline_frac": mul_bytes / (mul_ms * 1e-3) / 1e9 / peak,
                           "note": "per GPU; NTT / INTT batched (8 ciphertexts per launch pair), EvalMult+relin one ciphertext pair per call"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * 2 * l * N * 8, "d2h_bytes_per_step": B * 2 * l * N * 8,
                    "steps": e2e_steps, "matches_device_path": ok, "timed_regions_rot_per_s": [round(x, 1) for x in e2e_runs],
                    "GBps_each_way": e2e_val / world * 2 * l * N * 8 / 1e9, "pcie_duplex_probe_GBps_each_way": pcie_gbs,
                    "note": "overlapped fl_host_rotate_batch_async calls + one fl_sync per timed region; median of three regions; "
                            "bound by the host link (probe = plain pinned copies both ways at once, measured right after)"},
            "gpu_launches": launches_per_group * len(ins) * a.steps,
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if fwd:
            line["forward"] = fwd
        if sharded_blk:
            line["sharded_keyswitch"] = sharded_blk
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def forward_cpu_estimate(a, led, K, alpha):
    """Same-host CPU time of one forward, anchored on REAL runs of the CPU restatement: every operation type of the op ledger
    (EvalRotate, EvalMult + relinearisation, ct x pt, ct + ct, rescale) is timed by oracle/ckks_oracle.c (OpenMP, all host cores)
    at four limb counts of the forward's ring, and every ledger row (op @ l, count) is charged the time interpolated linearly in
    l for its own l.  `coverage` is the fraction of the ledger's algorithmic bytes whose operation type was timed this way
    (untimed: the plaintext-weighted sums of the encrypted E / F projection, and the host-side encodings, which the ledger does
    not carry at all) -- so the figure is still a lower bound, but no longer an extrapolation from key switches alone."""
    from oracle.oracle import Oracle, lib
    lib().orc_set_threads(host_threads())
    o = Oracle(logN=a.forward_logn, L=28, dnum=4)
    rng = np.random.default_rng(5)
    levels = (28, 21, 14, 7)

    def rand(l):
        return np.stack([rng.integers(0, int(o.moduli[m]), o.N, dtype=np.uint64) for m in range(l)])

    evk = rng.integers(0, 1 << 50, (o.dnum, 2, o.L + o.K, o.N), dtype=np.uint64)
    g = o.galois(1)

    def timed(fn, reps):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    sampled = {}
    for l in levels:
        ct, ct2, pt = np.stack([rand(l), rand(l)]), np.stack([rand(l), rand(l)]), rand(l)
        idx = list(range(l))
        sampled[l] = {
            "rotate": timed(lambda: o.rotate(ct, g, evk), 3),
            "mul_relin": timed(lambda: o.mul_relin(ct, ct2, evk), 2),
            "mul_plain": timed(lambda: o.mul_plain(ct, pt), 5),
            "add": timed(lambda: (o.add(ct[0], ct2[0], idx), o.add(ct[1], ct2[1], idx)), 5),
            "rescale": timed(lambda: (o.rescale(ct[0]), o.rescale(ct[1])), 2) if l >= 2 else 0.0,
        }
    fits = {}
    for op in ("rotate", "mul_relin", "mul_plain", "add", "rescale"):
        fits[op] = np.polyfit(np.array(levels, float), np.array([sampled[l][op] for l in levels], float), 1)
    total, per_op, timed_bytes, all_bytes, n_ops = 0.0, {}, 0.0, 0.0, 0
    for key, (n, nbytes) in led.items():
        op, _, l = key.partition("@")
        all_bytes += nbytes
        if op not in fits:
            continue
        t = n * max(0.0, float(np.polyval(fits[op], float(l))))
        total += t
        per_op[op] = per_op.get(op, 0.0) + t
        timed_bytes += nbytes
        n_ops += n
    cores = int(lib().orc_num_threads())
    return {"seconds_per_sample_lower_bound": float(total), "operations_charged": n_ops, "coverage": timed_bytes / max(all_bytes, 1.0),
            "seconds_by_operation": {k: round(v, 3) for k, v in per_op.items()}, "cores": cores, "kind": "port",
            "sampled_ms": {str(l): {k: round(v * 1e3, 3) for k, v in sampled[l].items()} for l in levels},
            "method": "oracle/ckks_oracle.c (OpenMP, all host cores) run for every ledger operation type at l = 28, 21, 14, 7 on the forward's ring; "
                      "each ledger row charged the time interpolated for its l; coverage = share of the ledger's algorithmic bytes timed; "
                      "OpenFHE-equivalent CPU restatement, not OpenFHE"}


def _quiet_stdout():
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)                       # the controller prints the reference's progress messages on stdout
    return devnull, saved


FORWARD_CONFIGS = [   # BASELINE.json configs 3 / 4 as SURVEY.md section 8(d) shapes them: name, classes, S, E/F projection under encryption
    ("config 3: IMDB-shaped (2 classes, S=256), client-projected E/F", 2, 256, False),
    ("config 3: IMDB-shaped (2 classes, S=256), E/F projection under encryption", 2, 256, True),
    ("config 4: BBC-shaped (5 classes, S=200), all bootstraps", 5, 200, False),
    ("config 4: 20NG-shaped (20 classes, S=256), all bootstraps", 20, 256, False),
]


def run_forward(a, local, rank, world, torch, dist):
    """BASELINE.json's first metric: encrypted Linformer forward, seconds/sample and samples/s, at the reference CKKS
    parameters (N=2^15, 28 limbs, 2^14 slots) through libflhost.so (FHEController + the main.cpp pipeline).
      * headline: SURVEY Config 1 shape (8 classes, S = --forward-rows = 200), each rank its own sample;
      * s129: the circuit's minimum S (round 1's figure);  configs: the Config 3 / 4 shapes;  n16: the commented-out ring;
      * batch: BASELINE config 5 -- a batch of samples sharded over the ranks (shard.my_units), one resident set of controllers per
        GPU in throughput mode (--forward-in-flight forwards at a time), logits gathered on rank 0 (shard.gather_logits);
        samples/s = samples of ALL ranks / max-over-ranks wall time."""
    import queue
    import tempfile
    from fhe_linformer_b200 import host, shard, synth
    root = tempfile.mkdtemp(prefix="flb200_bench_%d_" % rank)
    model = synth.make_model(n_classes=8)
    sample = synth.make_sample(model, a.forward_rows - 1, seed=20261018 + 1 + rank)
    dirs = synth.write_files(root, model, sample)
    in_flight = max(1, a.forward_in_flight)
    cache_gb = max(16, 120 // in_flight)      # block-cache cap per controller: a context parameter (fl_ctx_set_cache_bytes)
    devnull, saved = _quiet_stdout()
    try:
        t_k = time.perf_counter()
        fc = host.FHEController(device=local, root=root, cache_gb=cache_gb).generate(log_ring=0 if a.forward_logn == 15 else a.forward_logn)
        keygen_s = time.perf_counter() - t_k
        key_gb = fc.rotation_key_bytes() / 1e9
        fc.forward(dirs, dead_work=True)      # warm-up: mask / weight encodings, allocator pool (all keys were generated above)
        fc.ckks.ledger(True); fc.ckks.ledger_reset()
        runs = []
        for rep in range(3):                  # three timed samples: shared boxes show +-20 % run-to-run noise; the median is reported
            if rep == 1:
                led = fc.ckks.ledger_dump(); fc.ckks.ledger(False)
            t0 = time.perf_counter()
            logits, stages_i, S = fc.forward(dirs, dead_work=True)
            runs.append((time.perf_counter() - t0, stages_i))
        runs.sort(key=lambda r: r[0])
        dt, stages = runs[1]
        fc.forward(dirs, dead_work=False)     # warm-up of the lean variant (different batch shapes)
        lean_runs = []
        for rep in range(3):
            t1 = time.perf_counter()
            fc.forward(dirs, dead_work=False)
            lean_runs.append(time.perf_counter() - t1)
        lean = sorted(lean_runs)[1]
        # packed mode: the same network with the FFN on 128 rows per ciphertext through BSGS diagonal products (north star)
        fc.set_option("packed_keys", 1)
        fc.forward(dirs, packed=True); fc.forward(dirs, packed=True)      # warm-up: plans, plaintext diagonals at their levels
        fc.ckks.ledger(True); fc.ckks.ledger_reset()
        packed_runs = []
        for rep in range(3):
            if rep == 1:
                led_p = fc.ckks.ledger_dump(); fc.ckks.ledger(False)
            tp0 = time.perf_counter()
            logits_p, stages_p, _ = fc.forward(dirs, packed=True)
            packed_runs.append(time.perf_counter() - tp0)
        packed = {"seconds_per_sample": sorted(packed_runs)[1], "timed_samples_s": [round(x, 4) for x in packed_runs],
                  "rotations": sum(n for k, (n, _) in led_p.items() if k.startswith("rotate@")),
                  "max_logit_difference_to_faithful": float(np.abs(logits_p - logits).max()), "predicted_class": int(np.argmax(logits_p)),
                  "stage_seconds": stages_p, "rotation_keys_GB": fc.rotation_key_bytes() / 1e9,
                  "note": "LinformerForward::set_packed: rows stay wrapped-expanded (128 per ciphertext) between the two affines, every 128 x 128 FFN "
                          "block is one FHEController::packed_linear (16 x 8 baby/giant steps, double hoisting); same logits up to CKKS noise"}
        # packed + lean: only what the logits read (first half through the FFN, second affine + CLS mask folded into W2, 3 bootstraps)
        fc.forward(dirs, packed=True, dead_work=False); fc.forward(dirs, packed=True, dead_work=False)
        pl_runs = []
        for rep in range(3):
            tp0 = time.perf_counter()
            logits_pl, stages_pl, _ = fc.forward(dirs, packed=True, dead_work=False)
            pl_runs.append(time.perf_counter() - tp0)
        packed["lean"] = {"seconds_per_sample": sorted(pl_runs)[1], "timed_samples_s": [round(x, 4) for x in pl_runs],
                          "max_logit_difference_to_faithful": float(np.abs(logits_pl - logits).max()), "predicted_class": int(np.argmax(logits_pl)),
                          "stage_seconds": stages_pl,
                          "note": "packed mode minus the operations whose results the logits never read: second half of the rows skipped after the attention block, "
                                  "second affine and CLS column mask folded into the W2 diagonals, no refresh between GELU and the pooler"}

        def one_shape(classes, S_rows, encp, seed):
            m = synth.make_model(n_classes=classes)
            d = synth.write_files(tempfile.mkdtemp(prefix="flb200_cfg_%d_" % rank), m, synth.make_sample(m, S_rows - 1, seed=seed))
            fc.forward(d, dead_work=True, encrypted_projection=encp)
            ts = []
            for _ in range(2):
                tq = time.perf_counter()
                lg, _, _ = fc.forward(d, dead_work=True, encrypted_projection=encp)
                ts.append(time.perf_counter() - tq)
            return min(ts), int(np.argmax(lg[:classes]))

        s129 = None
        if a.forward_rows != 129:
            t129, c129 = one_shape(8, 129, False, 20261018 + 1 + rank)
            s129 = {"seconds_per_sample": t129, "rows_S": 129, "predicted_class": c129, "note": "the circuit's minimum S (round 1's headline shape)"}
        cfgs = []
        if not a.no_forward_configs and rank == 0:
            for name, ncls, S_rows, encp in FORWARD_CONFIGS:
                tc, cc = one_shape(ncls, S_rows, encp, 3000 + S_rows + ncls)
                cfgs.append({"name": name, "classes": ncls, "rows_S": S_rows, "encrypted_projection": encp, "seconds_per_sample": tc, "predicted_class": cc})

        # ---- BASELINE config 5: a batch of samples, sample-parallel over the ranks, throughput mode on every GPU ----
        total = a.samples if a.samples > 0 else 8 * world
        mine = list(shard.my_units(total, rank, world))
        wdir = dirs["weights"]
        jobs = []
        for i in mine:
            sd = os.path.join(root, "batch", str(i))
            smp = synth.make_sample(model, a.forward_rows - 1, seed=20261018 + 1000 + i)
            synth.write_sample_files(os.path.join(sd, "input"), os.path.join(sd, "tokens"), smp)
            jobs.append((i, {"weights": wdir, "input": os.path.join(sd, "input"), "tokens": os.path.join(sd, "tokens")}))
        ctl = [fc]
        for t in range(1, min(in_flight, max(1, len(jobs)))):
            root_t = tempfile.mkdtemp(prefix="flb200_bench_%d_%d_" % (rank, t))
            fc_t = host.FHEController(device=local, root=root_t, cache_gb=cache_gb).generate()
            fc_t.set_option("packed_keys", 1)
            fc_t.forward(dirs, dead_work=True)            # warm-up of this controller (encodings, pool)
            fc_t.forward(dirs, packed=True); fc_t.forward(dirs, packed=True)
            ctl.append(fc_t)

        def run_batch(use_packed, group=1, lean=False, controllers=None):
            # group > 1: `group` samples per call (flh_forward_many: every ciphertext of the forward carries one element per sample)
            q = queue.Queue()
            for k in range(0, len(jobs), group):
                q.put(jobs[k:k + group])
            res = {}
            failed = []

            def work(c):
                try:
                    work_(c)
                except Exception as exc:      # a worker must not die silently: the timing would cover fewer samples
                    failed.append(exc)

            def work_(c):
                while True:
                    try:
                        part = q.get_nowait()
                    except queue.Empty:
                        return
                    if group > 1:
                        z, _ = c.forward_many([d for _, d in part], dead_work=not lean, packed=True)
                        for (i, _), row in zip(part, z):
                            res[i] = row
                    else:
                        i, d = part[0]
                        res[i] = c.forward(d, packed=True, dead_work=not lean)[0] if use_packed else c.forward(d, dead_work=True)[0]

            if world > 1:
                dist.barrier()
            th = [threading.Thread(target=work, args=(c,)) for c in (controllers or ctl)]
            tb = time.perf_counter()
            for x in th: x.start()
            for x in th: x.join()
            if failed:
                raise failed[0]
            return time.perf_counter() - tb, res

        n_ctl = len(ctl)
        batch_dt, got = run_batch(False)
        batch_packed_dt, got_packed = run_batch(True)
        batch_agree = float(max(np.abs(got[i] - got_packed[i]).max() for i in got)) if got else 0.0
        for c in ctl[1:]:
            c.close()
        # several samples per call: one controller (a call already spreads every launch over its samples)
        per_call = max(1, min(a.samples_per_call, len(jobs)))
        run_batch(True, per_call, controllers=[fc]); run_batch(True, per_call, lean=True, controllers=[fc])            # warm-up of the batched shapes
        batch_many_dt, got_many = run_batch(True, per_call, controllers=[fc])
        batch_many_lean_dt, got_many_lean = run_batch(True, per_call, lean=True, controllers=[fc])
        many_agree = float(max(max(np.abs(got[i] - got_many[i]).max(), np.abs(got[i] - got_many_lean[i]).max()) for i in got)) if got else 0.0
        fc.close()
        n16 = None
        if a.forward_logn == 15 and not a.no_forward_n16 and rank == 0:
            # the same sample at the ring the reference leaves commented out (F.cpp:12): 2^14 slots become sparse packing
            fc16 = host.FHEController(device=local, root=root).generate(log_ring=16)
            fc16.forward(dirs, dead_work=True)
            t2 = time.perf_counter()
            logits16, _, _ = fc16.forward(dirs, dead_work=True)
            n16 = {"seconds_per_sample": time.perf_counter() - t2, "ring": "N=2^16, 28 limbs, dnum 4, 2^14 slots (sparse packing)",
                   "predicted_class": int(np.argmax(logits16)), "max_logit_difference_to_N15": float(np.abs(logits16 - logits).max())}
            fc16.close()
    finally:
        os.dup2(saved, 1)
        os.close(devnull); os.close(saved)
    t = torch.tensor([dt, lean, batch_dt, batch_packed_dt, packed["seconds_per_sample"], batch_agree, batch_many_dt, batch_many_lean_dt, many_agree],
                     device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, lean, batch_worst, batch_packed_worst = float(t[0].item()), float(t[1].item()), float(t[2].item()), float(t[3].item())
    packed["seconds_per_sample"] = float(t[4].item())
    packed["samples_per_s"] = world / packed["seconds_per_sample"]
    # the only data that leaves a rank: its samples' logits (padded to the largest share)
    per_rank = max(len(shard.my_units(total, r, world)) for r in range(world))
    mat = np.full((per_rank, 20), np.nan)
    for k, i in enumerate(mine):
        mat[k] = got[i]
    gathered = shard.gather_logits(mat, device="cuda")
    classes = [int(np.argmax(row[:8])) for m_ in gathered for row in m_ if not np.isnan(row[0])]
    batch = {"samples": total, "samples_per_gpu": [len(shard.my_units(total, r, world)) for r in range(world)], "rows_S": a.forward_rows,
             "in_flight_per_gpu": n_ctl, "seconds": batch_worst, "samples_per_s": total / batch_worst, "logits_gathered": len(classes),
             "predicted_class_histogram": {str(c): classes.count(c) for c in sorted(set(classes))},
             "block_cache_GB_per_controller": cache_gb,
             "packed": {"seconds": batch_packed_worst, "samples_per_s": total / batch_packed_worst, "max_logit_difference_to_faithful": float(t[5].item())},
             "packed_many": {"samples_per_call": per_call, "seconds": float(t[6].item()), "samples_per_s": total / float(t[6].item()),
                             "lean_seconds": float(t[7].item()), "lean_samples_per_s": total / float(t[7].item()),
                             "max_logit_difference_to_faithful": float(t[8].item()),
                             "note": "flh_forward_many: several samples per packed forward, every ciphertext one element per sample (ciphertext-parallel over "
                                     "samples inside each launch); lean = only what the logits read"},
             "note": "BASELINE config 5 (named size: 256 samples, --samples 256): samples sharded over the ranks (shard.my_units), resident "
                     "controllers in throughput mode, logits gathered on rank 0; samples/s = all samples / max-over-ranks wall time"}
    rot = sum(n for k, (n, _) in led.items() if k.startswith("rotate@"))
    alg = sum(b for _, b in led.values())
    return {"seconds_per_sample": dt, "samples_per_s": world / dt, "rows_S": S, "ring": "N=2^%d, 28 limbs, dnum 4, 2^14 slots" % a.forward_logn,
            "shape": "SURVEY.md Config 1: 8 classes, S = %d rows (CLS + %d tokens), d = 128, k = 32, FFN 512" % (S, S - 1),
            "rotations": rot, "algorithmic_GB": alg / 1e9, "achieved_GBps": alg / 1e9 / dt, "stage_seconds": stages,
            "lean_seconds_per_sample": lean, "predicted_class": int(np.argmax(logits)),
            "timed_samples_s": [round(r[0], 4) for r in runs], "lean_timed_samples_s": [round(x, 4) for x in sorted(lean_runs)],
            "packed": packed, "s129": s129, "configs": cfgs, "n16": n16, "batch": batch,
            "keys": {"rotation_keys_GB": key_gb, "context_and_key_generation_s": keygen_s,
                     "note": "every rotation key (listed + hoisted-ladder + tree indices) is generated before the first forward; none inside a timed region"},
            "_ledger": led,
            "note": "text files -> encode/encrypt -> encoder1 -> pooler -> classifier -> decrypt, wall clock incl. host encode; "
                    "lean = same logits without the operations main.cpp issues but never reads"}


def main():
    a = parse()
    claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
