"""Scheme-level ctypes front-end (keys, encode/encrypt, leveled ops, polynomial evaluation, bootstrap)
over include/fl_ckks.h.  Mirrors what FHEController calls on OpenFHE's CryptoContext."""
import ctypes as C
import math

import numpy as np

from .capi import Engine, _ptr

vp, ci, u32, u64, dbl = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double
CHEB_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)

_SIGS = {
    "fl_keygen": (ci, [vp, u64]), "fl_keygen_seeded": (ci, [vp, u64]), "fl_gen_mult_key": (ci, [vp]), "fl_gen_rot_keys": (ci, [vp, vp, ci]), "fl_gen_conj_key": (ci, [vp]),
    "fl_keys_clear": (ci, [vp, ci]), "fl_num_rot_keys": (ci, [vp]),
    "fl_export_sk": (ci, [vp, vp]), "fl_export_pk": (ci, [vp, vp]), "fl_export_evk": (ci, [vp, u32, vp]),
    "fl_import_keys": (ci, [vp, vp, vp]), "fl_import_evk": (ci, [vp, u32, vp]),
    "fl_keys_save": (ci, [vp, C.c_char_p]), "fl_keys_load": (ci, [vp, C.c_char_p]),
    "fl_encode": (ci, [vp, vp, vp, ci, ci, ci, C.POINTER(vp)]),
    "fl_encode_many": (ci, [vp, vp, ci, ci, ci, ci, C.POINTER(vp)]),
    "fl_encrypt_values_many": (ci, [vp, vp, ci, ci, ci, ci, C.POINTER(vp)]),
    "fl_encrypt": (ci, [vp, vp, C.POINTER(vp)]), "fl_encrypt_seeded": (ci, [vp, vp, u64, C.POINTER(vp)]),
    "fl_encrypt_many": (ci, [vp, vp, ci, C.POINTER(vp)]),
    "fl_decrypt": (ci, [vp, vp, vp, vp, ci]), "fl_decode": (ci, [vp, vp, vp, vp, ci]),
    "fl_add": (ci, [vp, vp, vp, C.POINTER(vp)]), "fl_sub": (ci, [vp, vp, vp, C.POINTER(vp)]), "fl_mul": (ci, [vp, vp, vp, C.POINTER(vp)]),
    "fl_add_many": (ci, [vp, vp, ci, C.POINTER(vp)]), "fl_mul_many": (ci, [vp, vp, ci, C.POINTER(vp)]),
    "fl_add_const": (ci, [vp, vp, dbl, C.POINTER(vp)]), "fl_mul_const": (ci, [vp, vp, dbl, C.POINTER(vp)]),
    "fl_rotate": (ci, [vp, vp, ci, C.POINTER(vp)]), "fl_rotsum": (ci, [vp, vp, ci, ci, C.POINTER(vp)]), "fl_conjugate": (ci, [vp, vp, C.POINTER(vp)]), "fl_rescale": (ci, [vp, vp, C.POINTER(vp)]),
    "fl_eval_poly": (ci, [vp, vp, vp, ci, C.POINTER(vp)]),
    "fl_eval_chebyshev": (ci, [vp, vp, vp, ci, dbl, dbl, C.POINTER(vp)]),
    "fl_chebyshev_coefficients": (ci, [CHEB_FN, vp, dbl, dbl, ci, vp]),
    "fl_bootstrap_setup": (ci, [vp, ci, ci, ci]), "fl_bootstrap_keygen": (ci, [vp, ci]), "fl_bootstrap": (ci, [vp, vp, C.POINTER(vp)]),
    "fl_lt_create": (ci, [vp, vp, ci, vp, vp, ci, ci, ci, C.POINTER(vp)]), "fl_lt_rotations": (ci, [vp, vp, vp, ci]),
    "fl_lt_shape": (ci, [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]),
    "fl_lt_apply": (ci, [vp, vp, vp, C.POINTER(vp)]), "fl_lt_apply_plain": (ci, [vp, vp, vp, C.POINTER(vp)]), "fl_lt_free": (None, [vp]),
    "fl_batch_pack": (ci, [vp, vp, ci, C.POINTER(vp)]), "fl_batch_slice": (ci, [vp, vp, ci, C.POINTER(vp)]), "fl_batch_range": (ci, [vp, vp, ci, ci, C.POINTER(vp)]), "fl_elem_batch": (ci, [vp]),
    "fl_elem_level": (ci, [vp]), "fl_elem_limbs": (ci, [vp]), "fl_elem_deg": (ci, [vp]), "fl_elem_slots": (ci, [vp]), "fl_elem_ncomp": (ci, [vp]),
    "fl_elem_scale": (dbl, [vp]), "fl_elem_clone": (ci, [vp, vp, C.POINTER(vp)]), "fl_elem_free": (None, [vp]),
    "fl_elem_export": (ci, [vp, vp, vp]), "fl_elem_import": (ci, [vp, vp, ci, ci, ci, dbl, ci, C.POINTER(vp)]),
    "fl_elem_save": (ci, [vp, vp, C.c_char_p]), "fl_elem_load": (ci, [vp, C.c_char_p, C.POINTER(vp)]),
    "fl_prof_enable": (ci, [vp, ci]), "fl_prof_dump": (ci, [vp, C.c_char_p, C.c_size_t]),
    "fl_ledger_enable": (ci, [vp, ci]), "fl_ledger_reset": (ci, [vp]), "fl_ledger_dump": (ci, [vp, C.c_char_p, C.c_size_t]),
}


class LinearTransform:
    """BSGS plan of a diagonal matrix (fl_lt*): rotations() lists the keys it needs, apply() is the double-hoisted product."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __del__(self):
        try:
            if self.h and self.ctx.h is not None:
                self.ctx.lib.fl_lt_free(self.h)
        except Exception:
            pass
        self.h = None

    def rotations(self):
        out = np.zeros(64, np.int32)
        n = self.ctx.lib.fl_lt_rotations(self.ctx.h, self.h, _ptr(out), 64)
        return [int(k) for k in out[:n]]

    @property
    def shape(self):
        a, b, g, d = ci(), ci(), ci(), ci()
        self.ctx.lib.fl_lt_shape(self.h, C.byref(a), C.byref(b), C.byref(g), C.byref(d))
        return {"n1": a.value, "n2": b.value, "stride": g.value, "diagonals": d.value}

    def apply(self, ct): return self.ctx._out(self.ctx.lib.fl_lt_apply, self.h, ct.h)
    def apply_plain(self, ct): return self.ctx._out(self.ctx.lib.fl_lt_apply_plain, self.h, ct.h)


class Elem:
    """Ciphertext or plaintext handle (fl_elem*)."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __del__(self):
        try:
            if self.h and self.ctx.h is not None:
                self.ctx.lib.fl_elem_free(self.h)
        except Exception:
            pass
        self.h = None

    level = property(lambda s: s.ctx.lib.fl_elem_level(s.h))       # Ciphertext::GetLevel()
    limbs = property(lambda s: s.ctx.lib.fl_elem_limbs(s.h))
    deg = property(lambda s: s.ctx.lib.fl_elem_deg(s.h))
    slots = property(lambda s: s.ctx.lib.fl_elem_slots(s.h))
    ncomp = property(lambda s: s.ctx.lib.fl_elem_ncomp(s.h))
    scale = property(lambda s: s.ctx.lib.fl_elem_scale(s.h))

    def GetLevel(self): return self.level
    def GetSlots(self): return self.slots
    def Clone(self): return self.ctx.clone(self)

    def export(self):
        out = np.empty((self.ncomp, self.limbs, self.ctx.N), np.uint64)
        self.ctx._ck(self.ctx.lib.fl_elem_export(self.ctx.h, self.h, _ptr(out)))
        return out


class CKKS(Engine):
    def __init__(self, device=0, **params):
        super().__init__(device=device, **params)
        for name, (res, args) in _SIGS.items():
            f = getattr(self.lib, name)
            f.restype, f.argtypes = res, args

    def _out(self, fn, *args):
        h = vp()
        self._ck(fn(self.h, *args, C.byref(h)))
        return Elem(self, h)

    # keys
    def keygen(self, seed=0):
        """seed 0: operating-system randomness (ChaCha20); any other value: reproducible TEST keys (fl_keygen_seeded)."""
        self._ck(self.lib.fl_keygen_seeded(self.h, seed) if seed else self.lib.fl_keygen(self.h, 0))
    def gen_mult_key(self): self._ck(self.lib.fl_gen_mult_key(self.h))
    def gen_rot_keys(self, idx):
        a = np.ascontiguousarray(idx, np.int32); self._ck(self.lib.fl_gen_rot_keys(self.h, _ptr(a), len(a)))
    def gen_conj_key(self): self._ck(self.lib.fl_gen_conj_key(self.h))
    def clear_keys(self, kind): self._ck(self.lib.fl_keys_clear(self.h, kind))
    def num_rot_keys(self): return self.lib.fl_num_rot_keys(self.h)
    def export_sk(self):
        o = np.empty((self.L + self.K, self.N), np.uint64); self._ck(self.lib.fl_export_sk(self.h, _ptr(o))); return o
    def export_pk(self):
        o = np.empty((2, self.L, self.N), np.uint64); self._ck(self.lib.fl_export_pk(self.h, _ptr(o))); return o
    def export_evk(self, g):
        o = np.empty((self.dnum, 2, self.L + self.K, self.N), np.uint64); self._ck(self.lib.fl_export_evk(self.h, g, _ptr(o))); return o
    def import_keys(self, sk=None, pk=None):
        self._ck(self.lib.fl_import_keys(self.h, _ptr(np.ascontiguousarray(sk)) if sk is not None else None,
                                         _ptr(np.ascontiguousarray(pk)) if pk is not None else None))
    def import_evk(self, g, evk): self._ck(self.lib.fl_import_evk(self.h, g, _ptr(np.ascontiguousarray(evk))))
    def save_keys(self, path): self._ck(self.lib.fl_keys_save(self.h, path.encode()))
    def load_keys(self, path): self._ck(self.lib.fl_keys_load(self.h, path.encode()))

    # encode / encrypt
    def encode(self, vals, level=0, slots=None):
        v = np.asarray(vals)
        slots = slots or self.N // 2
        re = np.ascontiguousarray(v.real, np.float64)
        im = np.ascontiguousarray(v.imag, np.float64) if np.iscomplexobj(v) else None
        return self._out(self.lib.fl_encode, _ptr(re), _ptr(im) if im is not None else None, len(re), level, slots)

    def encode_many(self, rows, level=0, slots=None):
        """Several real vectors of one length as ONE batched plaintext (fl_encode_many); unpack() gives the elements."""
        m = np.ascontiguousarray(np.asarray(rows, np.float64))
        if m.ndim != 2: raise ValueError("encode_many: a 2-D array of real rows expected")
        slots = slots or self.N // 2
        return self._out(self.lib.fl_encode_many, _ptr(m), m.shape[0], m.shape[1], level, slots)

    def encrypt_values_many(self, rows, level=0, slots=None):
        """Encode + encrypt several real vectors in one pass (fl_encrypt_values_many); returns the list of ciphertexts."""
        m = np.ascontiguousarray(np.asarray(rows, np.float64))
        if m.ndim != 2: raise ValueError("encrypt_values_many: a 2-D array of real rows expected")
        return self.unpack(self._out(self.lib.fl_encrypt_values_many, _ptr(m), m.shape[0], m.shape[1], level, slots or self.N // 2))

    def encrypt(self, x, level=0, slots=None, seed=None):
        p = x if isinstance(x, Elem) else self.encode(x, level, slots)
        if seed is None:
            return self._out(self.lib.fl_encrypt, p.h)
        return self._out(self.lib.fl_encrypt_seeded, p.h, seed)

    def encrypt_many(self, plaintexts):
        """Encrypt of several plaintexts of one level in one batched call; returns the list of ciphertexts."""
        if isinstance(plaintexts, Elem): plaintexts = [plaintexts]          # one batched plaintext (encode_many)
        arr = (vp * len(plaintexts))(*[p.h for p in plaintexts])
        return self.unpack(self._out(self.lib.fl_encrypt_many, arr, len(plaintexts)))

    def decrypt(self, ct, slots=None, complex_out=False):
        slots = slots or ct.slots
        re = np.empty(slots, np.float64); im = np.empty(slots, np.float64)
        self._ck(self.lib.fl_decrypt(self.h, ct.h, _ptr(re), _ptr(im), slots))
        return re + 1j * im if complex_out else re

    def decode(self, pt, slots=None):
        slots = slots or pt.slots
        re = np.empty(slots, np.float64); im = np.empty(slots, np.float64)
        self._ck(self.lib.fl_decode(self.h, pt.h, _ptr(re), _ptr(im), slots))
        return re + 1j * im

    # ops
    def add(self, a, b):
        if isinstance(b, (int, float)): return self._out(self.lib.fl_add_const, a.h, float(b))
        return self._out(self.lib.fl_add, a.h, b.h)
    def sub(self, a, b): return self._out(self.lib.fl_sub, a.h, b.h)
    def mult(self, a, b):
        if isinstance(b, (int, float)): return self._out(self.lib.fl_mul_const, a.h, float(b))
        return self._out(self.lib.fl_mul, a.h, b.h)
    def _many(self, fn, v):
        arr = (vp * len(v))(*[e.h for e in v]); return self._out(fn, arr, len(v))
    def add_many(self, v): return self._many(self.lib.fl_add_many, v)
    def mult_many(self, v): return self._many(self.lib.fl_mul_many, v)
    def pack(self, v): return self._many(self.lib.fl_batch_pack, v)                 # ciphertexts of equal level / scale as one batched operand
    def unpack(self, b): return [self._out(self.lib.fl_batch_slice, b.h, i) for i in range(self.lib.fl_elem_batch(b.h))]
    def rotate(self, a, k): return self._out(self.lib.fl_rotate, a.h, int(k))
    def rotsum(self, a, steps, stride):
        """r <- r + rot(r, stride 2^i), i < steps (FHEController::rotsum / repeat); hoisted groups only where their extra keys exist."""
        return self._out(self.lib.fl_rotsum, a.h, int(steps), int(stride))
    def conjugate(self, a): return self._out(self.lib.fl_conjugate, a.h)
    def rescale(self, a): return self._out(self.lib.fl_rescale, a.h)
    def clone(self, a): return self._out(self.lib.fl_elem_clone, a.h)
    def eval_poly(self, a, coeffs):
        c = np.ascontiguousarray(coeffs, np.float64); return self._out(self.lib.fl_eval_poly, a.h, _ptr(c), len(c))

    def chebyshev_coefficients(self, f, a, b, degree):
        out = np.empty(degree + 1, np.float64)
        cb = CHEB_FN(lambda x, _u: float(f(x)))
        self._ck(self.lib.fl_chebyshev_coefficients(cb, None, a, b, degree, _ptr(out)))
        return out

    def eval_chebyshev(self, x, coeffs, a, b):
        c = np.ascontiguousarray(coeffs, np.float64)
        return self._out(self.lib.fl_eval_chebyshev, x.h, _ptr(c), len(c), float(a), float(b))

    def eval_chebyshev_function(self, f, x, a, b, degree):   # EvalChebyshevFunction(f, ct, a, b, degree)
        return self.eval_chebyshev(x, self.chebyshev_coefficients(f, a, b, degree), a, b)

    # BSGS diagonal ct x pt matrix product (fl_lt_*)
    def linear_transform(self, diagonals, slots=None, level=-1, max_baby=0):
        """diagonals: {shift: vector[slots]} with (M v)[p] = sum_d diag_d[p] v[(p + d) mod slots]; returns a LinearTransform plan."""
        slots = slots or self.N // 2
        shifts = np.ascontiguousarray(sorted(diagonals), np.int32)
        d = np.stack([np.asarray(diagonals[int(k)], np.complex128) for k in shifts])
        re, im = np.ascontiguousarray(d.real), np.ascontiguousarray(d.imag)
        h = vp()
        self._ck(self.lib.fl_lt_create(self.h, _ptr(shifts), len(shifts), _ptr(re), _ptr(im), slots, level, max_baby, C.byref(h)))
        return LinearTransform(self, h)

    def bootstrap_setup(self, budget=(3, 3), slots=None): self._ck(self.lib.fl_bootstrap_setup(self.h, budget[0], budget[1], slots or self.N // 2))
    def bootstrap_keygen(self, slots=None): self._ck(self.lib.fl_bootstrap_keygen(self.h, slots or self.N // 2))
    def bootstrap(self, a): return self._out(self.lib.fl_bootstrap, a.h)

    def import_elem(self, arr, deg, scale, slots):
        a = np.ascontiguousarray(arr, np.uint64)
        return self._out(self.lib.fl_elem_import, _ptr(a), a.shape[0], a.shape[1], deg, float(scale), slots)
    def save(self, a, path): self._ck(self.lib.fl_elem_save(self.h, a.h, path.encode()))
    def load(self, path): return self._out(self.lib.fl_elem_load, path.encode())

    def prof(self, on=True): self.lib.fl_prof_enable(self.h, 1 if on else 0)
    def prof_dump(self):
        """{entry point: (calls, gpu ms, host ms)} since the last dump."""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.fl_prof_dump(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            k, n, g, h = line.split()
            out[k] = (int(float(n)), float(g), float(h))
        return out

    def ledger(self, on=True): self.lib.fl_ledger_enable(self.h, 1 if on else 0)
    def ledger_reset(self): self.lib.fl_ledger_reset(self.h)
    def ledger_dump(self):
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.fl_ledger_dump(self.h, buf, len(buf)))
        rows = {}
        for line in buf.value.decode().splitlines():
            k, n, b = line.split()
            rows[k] = (int(n), float(b))
        return rows
