"""Limb-sharded hybrid key switch over NVLink (the optional multi-GPU mode of SURVEY.md section 8(e)).

One EvalRotate is split across the G ranks of a torch.distributed group by LIMB: rank r owns a contiguous range of the Q
limbs and a range of the K special limbs of the extended basis Q_l u P.  Every stage of the key switch is limb-wise
independent except the two base conversions, so a rank

  1. takes the digits to coefficient form (all l limbs: cheap, recomputed by every rank instead of exchanged),
  2. extends them (ModUp + NTT) to ITS limbs only and multiplies by ITS limbs of the evaluation key
     (so the key stream, the largest of a key switch, is split G ways),
  3. takes its special limbs of the two accumulators to coefficient form,
     --- exchange 1: the 2 K special limbs (7.3 MB at N = 2^16) are combined with one all-reduce over disjoint supports ---
  4. converts them down (ModDown) to ITS Q limbs, transforms, subtracts, scales and applies the automorphism,
     --- exchange 2: the 2 l result limbs (29 MB) are combined the same way, every rank ends with the full ciphertext ---

The per-limb kernels are the single-GPU ones restricted to a limb range (fl_raw_ks_* in include/fl_ckks.h), so the result is
bit-identical to the single-GPU rotation.  `LocalComm` runs the ranks one after the other on one GPU (for tests);
`DistComm` is NCCL (or gloo) through torch.distributed.  Collectives and the torch glue run on the engine's stream.
"""
import numpy as np
import torch

from . import shard


class DistComm:
    """Ranks are processes of a torch.distributed group (NCCL over NVLink on the GPU box)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)


class LocalComm:
    """All ranks in one process: tensors of the virtual ranks are summed in place (single-GPU emulation for tests)."""

    def __init__(self, world):
        self.world = world

    @staticmethod
    def all_reduce_sum_many(tensors):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total += t
        for t in tensors:
            t.copy_(total)


class _RankState:
    """Work buffers and limb ranges of one rank."""

    def __init__(self, eng, l, rank, world, device):
        self.e, self.l, self.rank, self.world = eng, l, rank, world
        K, N = eng.K, eng.N
        self.ext = l + K
        self.beta = -(-l // eng.alpha)
        q = shard.my_units(l, rank, world)
        p = shard.my_units(K, rank, world)
        self.q_first, self.q_count = (q.start, len(q))
        self.p_first, self.p_count = (p.start, len(p))
        i64 = dict(dtype=torch.int64, device=device)
        self.dco = torch.empty((l, N), **i64)
        self.up = torch.empty((self.beta, self.ext, N), **i64)
        self.acc = torch.zeros((2, self.ext, N), **i64)
        self.tq = torch.empty((2, l, N), **i64)
        self.out = torch.zeros((2, l, N), **i64)
        self.pbuf = torch.zeros((2, K, N), **i64)

    def ranges(self):
        """The rank's two ranges inside the extended basis: its Q limbs and its special limbs (shifted by l)."""
        return [(self.q_first, self.q_count), (self.l + self.p_first, self.p_count)]


def _p(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())


class ShardedKeySwitch:
    """EvalRotate with the limbs of Q_l u P split across ranks.  ct: int64 CUDA tensor [2][l][N] (the same on every rank),
    evk: DevBuf or tensor holding the full key (a rank only reads its limbs of it)."""

    def __init__(self, eng, l, comm, device=None):
        self.e, self.l, self.comm = eng, l, comm
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(eng.stream(), device=self.device)
        if isinstance(comm, LocalComm):
            self.states = [_RankState(eng, l, r, comm.world, self.device) for r in range(comm.world)]
        else:
            self.states = [_RankState(eng, l, comm.rank, comm.world, self.device)]

    def _ck(self, rc):
        self.e._ck(rc)

    def _front(self, st, ct, evk_ptr):
        lib, h, l = self.e.lib, self.e.h, self.l
        self._ck(lib.fl_raw_ks_digits(h, _p(st.dco), _p(ct[1]), l))
        for first, count in st.ranges():
            if count:
                self._ck(lib.fl_raw_ks_modup(h, _p(st.up), _p(st.dco), l, first, count))
                self._ck(lib.fl_raw_ks_inner(h, _p(st.acc), _p(st.up), _p(ct[1]), evk_ptr, l, first, count))
        if st.p_count:
            self._ck(lib.fl_raw_ks_pcoef(h, _p(st.acc), l, st.p_first, st.p_count))
        st.pbuf.zero_()
        if st.p_count:
            st.pbuf[:, st.p_first:st.p_first + st.p_count] = st.acc[:, l + st.p_first:l + st.p_first + st.p_count]

    def _back(self, st, ct, g):
        lib, h, l = self.e.lib, self.e.h, self.l
        st.acc[:, l:] = st.pbuf                       # every rank now holds all special limbs in coefficient form
        st.out.zero_()
        if st.q_count:
            self._ck(lib.fl_raw_ks_moddown(h, _p(st.out), _p(st.tq), _p(st.acc), l, st.q_first, st.q_count, _p(ct[0]), None, g))

    def rotate(self, ct, g, evk):
        """Returns the rotated ciphertext [2][l][N] (identical on every rank)."""
        evk_ptr = evk.ptr if hasattr(evk, "ptr") else _p(evk)
        with torch.cuda.stream(self.stream):
            for st in self.states:
                self._front(st, ct, evk_ptr)
            if isinstance(self.comm, LocalComm):
                LocalComm.all_reduce_sum_many([st.pbuf for st in self.states])
            else:
                self.comm.all_reduce_sum(self.states[0].pbuf)
            for st in self.states:
                self._back(st, ct, g)
            if isinstance(self.comm, LocalComm):
                LocalComm.all_reduce_sum_many([st.out for st in self.states])
            else:
                self.comm.all_reduce_sum(self.states[0].out)
        return self.states[0].out

    def exchanged_bytes(self):
        """Payload a rank contributes to the two exchanges of one rotation."""
        return (2 * self.e.K + 2 * self.l) * self.e.N * 8


def register_signatures(lib):
    import ctypes as C
    vp, ci, u32 = C.c_void_p, C.c_int, C.c_uint32
    for name, args in {
        "fl_raw_ks_digits": [vp, vp, vp, ci], "fl_raw_ks_modup": [vp, vp, vp, ci, ci, ci],
        "fl_raw_ks_inner": [vp, vp, vp, vp, vp, ci, ci, ci], "fl_raw_ks_pcoef": [vp, vp, ci, ci, ci],
        "fl_raw_ks_moddown": [vp, vp, vp, vp, ci, ci, ci, vp, vp, u32],
    }.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = ci, args


def to_tensor(a, device):
    """uint64 numpy array -> int64 CUDA tensor with the same bits."""
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(device)
