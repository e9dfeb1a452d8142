"""Limb-sharded hybrid key switch over NVLink (the optional multi-GPU mode of SURVEY.md section 8(e)).

One EvalRotate is split across the G ranks of a torch.distributed group by LIMB: rank r owns a contiguous range of the Q
limbs and a range of the K special limbs of the extended basis Q_l u P.  Every stage of the key switch is limb-wise
independent except the two base conversions, so a rank

  1. takes the digits to coefficient form -- either all l limbs itself (`gather_digits=False`: 28 small transforms, no traffic) or
     only its share, followed by
     --- exchange 0: ALL-GATHER of the l scaled INTT'd digit limbs (l N 8 bytes in total: 14.7 MB at N = 2^16) ---
  2. extends them (ModUp + NTT) to ITS limbs only and multiplies by ITS limbs of the evaluation key
     (so the key stream, the largest of a key switch, is split G ways),
  3. takes its special limbs of the two accumulators to coefficient form,
     --- exchange 1: ALL-GATHER of the 2 K special limbs (7.3 MB at N = 2^16; round 1 used an all-reduce over disjoint supports,
         which moves twice the bytes) ---
  4. converts them down (ModDown) to ITS Q limbs, transforms, subtracts, scales and applies the automorphism.

The result STAYS limb-sharded: rank r holds limbs [q_first, q_first + q_count) of both polynomials (round 1 all-reduced the
2 l result limbs, 29 MB, so that every rank ended with the full ciphertext; a chain of limb-wise operations does not need
that, and the next key switch gathers exactly what it needs in its exchange 0).  `gather_result()` assembles the full
ciphertext when a caller wants it (the tests do).

The per-limb kernels are the single-GPU ones restricted to a limb range (fl_raw_ks_* in include/fl_ckks.h), so the result is
bit-identical to the single-GPU rotation.  `LocalComm` runs the ranks one after the other on one GPU (for tests);
`DistComm` is NCCL (or gloo) through torch.distributed.  Collectives and the torch glue run on the engine's stream.
"""
import numpy as np
import torch

from . import shard


def share_sizes(total, world):
    """Limbs per rank of a contiguous block partition (shard.my_units), and the padded share all-gather uses."""
    sizes = [len(shard.my_units(total, r, world)) for r in range(world)]
    return sizes, max(sizes)


def assemble(gathered, total, world, axis):
    """gathered: [world][...padded share...] as returned by an all-gather of equally sized shares; returns the `total` limbs in
    order along `axis` (padding dropped).  Pure tensor logic: covered on CPU with gloo (tests/test_shard_gloo.py)."""
    sizes, _ = share_sizes(total, world)
    parts = [gathered[r].narrow(axis, 0, sizes[r]) for r in range(world) if sizes[r]]
    return torch.cat(parts, dim=axis)


class DistComm:
    """Ranks are processes of a torch.distributed group (NCCL over NVLink on the GPU box)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather(self, share):
        """share: this rank's (padded) block; returns a [world, *share.shape] tensor."""
        share = share.contiguous()
        out = torch.empty((self.world * share.shape[0],) + tuple(share.shape[1:]), dtype=share.dtype, device=share.device)   # rank-major concatenation
        self.dist.all_gather_into_tensor(out, share, group=self.group)
        return out.view((self.world,) + tuple(share.shape))


class LocalComm:
    """All ranks in one process (single-GPU emulation for tests): an all-gather is a stack of the virtual ranks' shares."""

    def __init__(self, world):
        self.world = world

    @staticmethod
    def all_gather_many(shares):
        return torch.stack([s.contiguous() for s in shares])


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
        _, self.q_pad = share_sizes(l, world)
        _, self.p_pad = share_sizes(K, world)
        i64 = dict(dtype=torch.int64, device=device)
        self.dco = torch.empty((l, N), **i64)
        self.up = torch.empty((self.beta, self.ext, N), **i64)
        self.acc = torch.zeros((2, self.ext, N), **i64)
        self.tq = torch.empty((2, l, N), **i64)
        self.out = torch.zeros((2, l, N), **i64)             # only limbs [q_first, q_first + q_count) are ever written
        self.dshare = torch.zeros((self.q_pad, N), **i64)    # exchange 0: this rank's digit limbs (padded to the largest share)
        self.pshare = torch.zeros((2, self.p_pad, N), **i64) # exchange 1: this rank's special limbs of both accumulators

    def ranges(self):
        """The rank's two ranges inside the extended basis: its Q limbs and its special limbs (shifted by l)."""
        return [(self.q_first, self.q_count), (self.l + self.p_first, self.p_count)]


def _p(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())


class ShardedKeySwitch:
    """EvalRotate with the limbs of Q_l u P split across ranks.  ct: int64 CUDA tensor [2][l][N] (the same on every rank),
    evk: DevBuf or tensor holding the full key (a rank only reads its limbs of it)."""

    def __init__(self, eng, l, comm, device=None, gather_digits=False):
        self.e, self.l, self.comm, self.gather_digits = eng, l, comm, gather_digits
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(eng.stream(), device=self.device)
        self.local = isinstance(comm, LocalComm)
        world = comm.world
        self.states = [_RankState(eng, l, r, world, self.device) for r in range(world)] if self.local else [_RankState(eng, l, comm.rank, world, self.device)]

    def _ck(self, rc):
        self.e._ck(rc)

    def _gather(self, shares):
        return LocalComm.all_gather_many(shares) if self.local else self.comm.all_gather(shares[0])

    def _digits(self, ct):
        lib, h, l = self.e.lib, self.e.h, self.l
        if not self.gather_digits:
            for st in self.states:
                self._ck(lib.fl_raw_ks_digits(h, _p(st.dco), _p(ct[1]), l))
            return
        for st in self.states:                                   # exchange 0: every rank transforms its share only
            if st.q_count:
                self._ck(lib.fl_raw_ks_digits_part(h, _p(st.dco), _p(ct[1]), l, st.q_first, st.q_count))
                st.dshare[:st.q_count] = st.dco[st.q_first:st.q_first + st.q_count]
        full = assemble(self._gather([st.dshare for st in self.states]), l, self.comm.world, 0)
        for st in self.states:
            st.dco.copy_(full)

    def _front(self, st, ct, evk_ptr):
        lib, h, l = self.e.lib, self.e.h, self.l
        for first, count in st.ranges():
            if count:
                self._ck(lib.fl_raw_ks_modup(h, _p(st.up), _p(st.dco), l, first, count))
                self._ck(lib.fl_raw_ks_inner(h, _p(st.acc), _p(st.up), _p(ct[1]), evk_ptr, l, first, count))
        if st.p_count:
            self._ck(lib.fl_raw_ks_pcoef(h, _p(st.acc), l, st.p_first, st.p_count))
            st.pshare[:, :st.p_count] = st.acc[:, l + st.p_first:l + st.p_first + st.p_count]

    def _back(self, st, ct, g):
        lib, h, l = self.e.lib, self.e.h, self.l
        if st.q_count:
            self._ck(lib.fl_raw_ks_moddown(h, _p(st.out), _p(st.tq), _p(st.acc), l, st.q_first, st.q_count, _p(ct[0]), None, g))

    def rotate(self, ct, g, evk):
        """Returns this rank's buffer [2][l][N] in which limbs [q_first, q_first + q_count) hold the rotated ciphertext (the
        result stays limb-sharded; see gather_result)."""
        evk_ptr = evk.ptr if hasattr(evk, "ptr") else _p(evk)
        K = self.e.K
        with torch.cuda.stream(self.stream):
            self._digits(ct)
            for st in self.states:
                self._front(st, ct, evk_ptr)
            # exchange 1: all ranks receive all special limbs of both accumulators (coefficient form)
            pall = assemble(self._gather([st.pshare for st in self.states]), K, self.comm.world, 1)
            for st in self.states:
                st.acc[:, self.l:] = pall
                self._back(st, ct, g)
        return self.states[0].out

    def gather_result(self):
        """The full rotated ciphertext [2][l][N] on every rank (an all-gather of the 2 l result limbs: only when a caller needs it)."""
        with torch.cuda.stream(self.stream):
            shares = []
            for st in self.states:
                s = torch.zeros((2, st.q_pad, self.e.N), dtype=torch.int64, device=self.device)
                s[:, :st.q_count] = st.out[:, st.q_first:st.q_first + st.q_count]
                shares.append(s)
            return assemble(self._gather(shares), self.l, self.comm.world, 1)

    def exchanged_bytes(self):
        """Bytes a rank RECEIVES in the exchanges of one rotation (all-gathers: everything but its own share)."""
        w = self.comm.world
        total = 2 * self.e.K * self.e.N * 8 + (self.l * self.e.N * 8 if self.gather_digits else 0)
        return total * (w - 1) // w


def register_signatures(lib):
    import ctypes as C
    vp, ci, u32 = C.c_void_p, C.c_int, C.c_uint32
    for name, args in {
        "fl_raw_ks_digits": [vp, vp, vp, ci], "fl_raw_ks_digits_part": [vp, vp, vp, ci, ci, ci], "fl_raw_ks_modup": [vp, vp, vp, ci, ci, ci],
        "fl_raw_ks_inner": [vp, vp, vp, vp, vp, ci, ci, ci], "fl_raw_ks_pcoef": [vp, vp, ci, ci, ci],
        "fl_raw_ks_moddown": [vp, vp, vp, vp, ci, ci, ci, vp, vp, u32],
    }.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = ci, args


def to_tensor(a, device):
    """uint64 numpy array -> int64 CUDA tensor with the same bits."""
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(device)
