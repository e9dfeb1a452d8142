"""ctypes loader for the CPU oracle (oracle/ckks_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package (fhe_linformer_b200).
Parity unpinned: see oracle/ckks_oracle.h.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libckks_oracle.so")

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "ckks_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        vp, ci, u64, i64, u32, dbl = C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_uint32, C.c_double
        sig = {
            "orc_create": (vp, [ci] * 6), "orc_destroy": (None, [vp]), "orc_info": (None, [vp, i32p]),
            "orc_moduli": (None, [vp, u64p]), "orc_roots": (None, [vp, u64p]), "orc_scale_factors": (None, [vp, f64p]),
            "orc_ntt": (None, [vp, u64p, i32p, ci]), "orc_intt": (None, [vp, u64p, i32p, ci]),
            "orc_add": (None, [vp, u64p, u64p, u64p, i32p, ci]), "orc_sub": (None, [vp, u64p, u64p, u64p, i32p, ci]),
            "orc_mul": (None, [vp, u64p, u64p, u64p, i32p, ci]),
            "orc_mul_scalar": (None, [vp, u64p, u64p, u64, i64, i32p, ci]),
            "orc_automorph_eval": (None, [vp, u64p, u64p, ci, u32]),
            "orc_automorph_coeff": (None, [vp, u64p, u64p, i32p, ci, u32]),
            "orc_galois_for_rotation": (u32, [vp, ci]), "orc_galois_conj": (u32, [vp]),
            "orc_rescale": (None, [vp, u64p, u64p, ci]),
            "orc_modup": (None, [vp, u64p, u64p, ci, ci]), "orc_moddown": (None, [vp, u64p, u64p, ci]),
            "orc_keyswitch": (None, [vp, u64p, u64p, u64p, u64p, ci]),
            "orc_rotate": (None, [vp, u64p, u64p, ci, u32, u64p]),
            "orc_mul_relin": (None, [vp, u64p, u64p, u64p, ci, u64p]),
            "orc_mul_plain": (None, [vp, u64p, u64p, u64p, ci]),
            "orc_sample_uniform": (None, [vp, u64, u64p, i32p, ci]),
            "orc_sample_ternary": (None, [vp, u64, i8p]), "orc_sample_sparse_ternary": (None, [vp, u64, ci, i8p]),
            "orc_sample_gauss": (None, [vp, u64, i8p]), "orc_small_to_eval": (None, [vp, u64p, i8p, i32p, ci]),
            "orc_gen_sk": (None, [vp, u64, ci, u64p]), "orc_gen_pk": (None, [vp, u64, u64p, u64p]),
            "orc_gen_evk": (None, [vp, u64, u64p, u64p, u64p]), "orc_gen_relin_key": (None, [vp, u64, u64p, u64p]),
            "orc_gen_galois_key": (None, [vp, u64, u64p, u32, u64p]),
            "orc_encode": (None, [vp, u64p, f64p, ci, dbl, ci]), "orc_encode_coeffs": (None, [vp, i64p, f64p, ci, dbl]),
            "orc_decode": (None, [vp, f64p, u64p, ci, dbl, ci]),
            "orc_encrypt": (None, [vp, u64, u64p, u64p, u64p, ci]), "orc_decrypt": (None, [vp, u64p, u64p, u64p, ci, ci]),
            "orc_num_threads": (ci, []), "orc_set_threads": (None, [ci]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name); f.restype = res; f.argtypes = args
        _lib = L
    return _lib


class Oracle:
    """Thin numpy front-end.  Polynomials are uint64 arrays [limbs, N]; ciphertexts [2, limbs, N]."""

    def __init__(self, logN=15, L=28, dnum=4, first_bits=55, scale_bits=52, aux_bits=60):
        self.l = lib()
        self.h = self.l.orc_create(logN, L, dnum, first_bits, scale_bits, aux_bits)
        info = np.zeros(8, np.int32); self.l.orc_info(self.h, info)
        self.logN, self.L, self.K, self.alpha, self.dnum = (int(x) for x in info[:5])
        self.N = 1 << self.logN
        self.moduli = np.zeros(self.L + self.K, np.uint64); self.l.orc_moduli(self.h, self.moduli)
        self.roots = np.zeros(self.L + self.K, np.uint64); self.l.orc_roots(self.h, self.roots)
        self.sf = np.zeros(self.L, np.float64); self.l.orc_scale_factors(self.h, self.sf)

    def __del__(self):
        try:
            self.l.orc_destroy(self.h)
        except Exception:
            pass

    # -- helpers
    def qidx(self, l):
        return np.arange(l, dtype=np.int32)

    def allidx(self):
        return np.arange(self.L + self.K, dtype=np.int32)

    def ext_idx(self, l):
        return np.concatenate([np.arange(l), self.L + np.arange(self.K)]).astype(np.int32)

    def _midx(self, a, midx):
        return self.qidx(a.shape[-2]) if midx is None else np.ascontiguousarray(midx, np.int32)

    # -- primitives
    def ntt(self, a, midx=None):
        a = np.ascontiguousarray(a.copy()); m = self._midx(a, midx); self.l.orc_ntt(self.h, a, m, len(m)); return a

    def intt(self, a, midx=None):
        a = np.ascontiguousarray(a.copy()); m = self._midx(a, midx); self.l.orc_intt(self.h, a, m, len(m)); return a

    def _bin(self, fn, a, b, midx):
        m = self._midx(a, midx); out = np.empty_like(a); fn(self.h, out, np.ascontiguousarray(a), np.ascontiguousarray(b), m, len(m)); return out

    def add(self, a, b, midx=None): return self._bin(self.l.orc_add, a, b, midx)
    def sub(self, a, b, midx=None): return self._bin(self.l.orc_sub, a, b, midx)
    def mul(self, a, b, midx=None): return self._bin(self.l.orc_mul, a, b, midx)

    def mul_scalar(self, a, s, midx=None):
        m = self._midx(a, midx); out = np.empty_like(a); s = int(s)
        lo = s & ((1 << 64) - 1); hi = s >> 64
        self.l.orc_mul_scalar(self.h, out, np.ascontiguousarray(a), lo, hi, m, len(m)); return out

    def galois(self, k): return int(self.l.orc_galois_for_rotation(self.h, int(k)))
    def galois_conj(self): return int(self.l.orc_galois_conj(self.h))

    def automorph_eval(self, a, g):
        out = np.empty_like(a); self.l.orc_automorph_eval(self.h, out, np.ascontiguousarray(a), a.shape[0], g); return out

    def automorph_coeff(self, a, g, midx=None):
        m = self._midx(a, midx); out = np.empty_like(a); self.l.orc_automorph_coeff(self.h, out, np.ascontiguousarray(a), m, len(m), g); return out

    def rescale(self, a):
        l = a.shape[0]; out = np.empty((l - 1, self.N), np.uint64); self.l.orc_rescale(self.h, out, np.ascontiguousarray(a), l); return out

    def modup(self, a, digit):
        l = a.shape[0]; out = np.empty((l + self.K, self.N), np.uint64); self.l.orc_modup(self.h, out, np.ascontiguousarray(a), l, digit); return out

    def moddown(self, a):
        l = a.shape[0] - self.K; out = np.empty((l, self.N), np.uint64); self.l.orc_moddown(self.h, out, np.ascontiguousarray(a), l); return out

    def keyswitch(self, a, evk):
        l = a.shape[0]; o0 = np.empty_like(a); o1 = np.empty_like(a)
        self.l.orc_keyswitch(self.h, o0, o1, np.ascontiguousarray(a), evk, l); return o0, o1

    def rotate(self, ct, g, evk):
        out = np.empty_like(ct); self.l.orc_rotate(self.h, out, np.ascontiguousarray(ct), ct.shape[1], g, evk); return out

    def mul_relin(self, a, b, evk):
        out = np.empty_like(a); self.l.orc_mul_relin(self.h, out, np.ascontiguousarray(a), np.ascontiguousarray(b), a.shape[1], evk); return out

    def mul_plain(self, ct, pt):
        out = np.empty_like(ct); self.l.orc_mul_plain(self.h, out, np.ascontiguousarray(ct), np.ascontiguousarray(pt), ct.shape[1]); return out

    # -- sampling / keys
    def sample_uniform(self, seed, midx):
        m = np.ascontiguousarray(midx, np.int32); out = np.empty((len(m), self.N), np.uint64)
        self.l.orc_sample_uniform(self.h, seed, out, m, len(m)); return out

    def sample_ternary(self, seed):
        o = np.empty(self.N, np.int8); self.l.orc_sample_ternary(self.h, seed, o); return o

    def sample_sparse(self, seed, h):
        o = np.empty(self.N, np.int8); self.l.orc_sample_sparse_ternary(self.h, seed, h, o); return o

    def sample_gauss(self, seed):
        o = np.empty(self.N, np.int8); self.l.orc_sample_gauss(self.h, seed, o); return o

    def small_to_eval(self, s, midx):
        m = np.ascontiguousarray(midx, np.int32); out = np.empty((len(m), self.N), np.uint64)
        self.l.orc_small_to_eval(self.h, out, np.ascontiguousarray(s, np.int8), m, len(m)); return out

    def gen_sk(self, seed, h=0):
        sk = np.empty((self.L + self.K, self.N), np.uint64); self.l.orc_gen_sk(self.h, seed, h, sk); return sk

    def gen_pk(self, seed, sk):
        pk = np.empty((2, self.L, self.N), np.uint64); self.l.orc_gen_pk(self.h, seed, sk, pk); return pk

    def _evk(self): return np.empty((self.dnum, 2, self.L + self.K, self.N), np.uint64)

    def gen_relin_key(self, seed, sk):
        e = self._evk(); self.l.orc_gen_relin_key(self.h, seed, sk, e); return e

    def gen_galois_key(self, seed, sk, g):
        e = self._evk(); self.l.orc_gen_galois_key(self.h, seed, sk, g, e); return e

    # -- encoding / encryption
    @staticmethod
    def _reim(vals):
        v = np.asarray(vals, np.complex128); out = np.empty(2 * len(v), np.float64); out[0::2] = v.real; out[1::2] = v.imag; return out

    def encode(self, vals, scale, l, slots=None):
        slots = slots or len(vals); out = np.empty((l, self.N), np.uint64)
        self.l.orc_encode(self.h, out, self._reim(vals), slots, float(scale), l); return out

    def encode_coeffs(self, vals, scale, slots=None):
        slots = slots or len(vals); out = np.empty(self.N, np.int64)
        self.l.orc_encode_coeffs(self.h, out, self._reim(vals), slots, float(scale)); return out

    def decode(self, poly, scale, slots):
        out = np.empty(2 * slots, np.float64)
        self.l.orc_decode(self.h, out, np.ascontiguousarray(poly), slots, float(scale), poly.shape[0]); return out[0::2] + 1j * out[1::2]

    def encrypt(self, seed, pt, pk):
        l = pt.shape[0]; ct = np.empty((2, l, self.N), np.uint64)
        self.l.orc_encrypt(self.h, seed, ct, np.ascontiguousarray(pt), pk, l); return ct

    def decrypt(self, ct, sk):
        l = ct.shape[1]; pt = np.empty((l, self.N), np.uint64)
        self.l.orc_decrypt(self.h, pt, np.ascontiguousarray(ct), sk, l, ct.shape[0]); return pt
