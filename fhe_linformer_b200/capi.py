"""ctypes binding of include/fl_ckks.h (the drop-in boundary).  Fails loudly when the CUDA library is
missing: there is deliberately no fallback path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libflckks.so")


class fl_params(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("logN", "L", "dnum", "first_bits", "scale_bits", "aux_bits", "sparse_h")]


# reference parameters, /root/reference/src/FHEController.cpp:6-35
ReferenceParams = dict(logN=15, L=28, dnum=4, first_bits=55, scale_bits=52, aux_bits=60, sparse_h=192)

_lib = None
u64p = C.POINTER(C.c_uint64)


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  fhe_linformer_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ci, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
    ip = C.POINTER(C.c_int)
    sig = {
        "fl_last_error": (C.c_char_p, []),
        "fl_ctx_create": (ci, [C.POINTER(fl_params), ci, C.POINTER(vp)]),
        "fl_ctx_destroy": (None, [vp]),
        "fl_ctx_info": (ci, [vp, ip]),
        "fl_ctx_moduli": (ci, [vp, vp]), "fl_ctx_roots": (ci, [vp, vp]), "fl_ctx_scale_factors": (ci, [vp, vp]),
        "fl_galois_for_rotation": (u32, [vp, ci]), "fl_galois_conj": (u32, [vp]),
        "fl_ctx_stream": (vp, [vp]), "fl_sync": (ci, [vp]),
        "fl_dev_alloc": (ci, [vp, sz, C.POINTER(vp)]), "fl_dev_free": (ci, [vp, vp]),
        "fl_dev_upload": (ci, [vp, vp, vp, sz]), "fl_dev_download": (ci, [vp, vp, vp, sz]),
        "fl_raw_ntt": (ci, [vp, vp, vp, ci]), "fl_raw_intt": (ci, [vp, vp, vp, ci]),
        "fl_raw_ntt_batch": (ci, [vp, vp, vp, ci, ci, ci]),
        "fl_raw_add": (ci, [vp, vp, vp, vp, vp, ci]), "fl_raw_sub": (ci, [vp, vp, vp, vp, vp, ci]),
        "fl_raw_mul": (ci, [vp, vp, vp, vp, vp, ci]),
        "fl_raw_automorph": (ci, [vp, vp, vp, ci, u32]),
        "fl_raw_rescale": (ci, [vp, vp, vp, ci, ci]),
        "fl_raw_modup": (ci, [vp, vp, vp, ci, ci]), "fl_raw_moddown": (ci, [vp, vp, vp, ci]),
        "fl_raw_keyswitch": (ci, [vp, vp, vp, vp, ci]),
        "fl_raw_rotate": (ci, [vp, vp, vp, ci, u32, vp]),
        "fl_raw_rotate_batch": (ci, [vp, vp, vp, ci, u32, vp, ci]),
        "fl_host_rotate_batch": (ci, [vp, vp, vp, ci, u32, vp, ci]), "fl_host_rotate_batch_async": (ci, [vp, vp, vp, ci, u32, vp, ci]),
        "fl_raw_mul_relin": (ci, [vp, vp, vp, vp, ci, vp]),
        "fl_raw_mul_plain": (ci, [vp, vp, vp, vp, ci]),
        "fl_host_ntt": (ci, [vp, vp, ci, ci]),
        "fl_host_rotate": (ci, [vp, vp, vp, ci, u32, vp]),
        "fl_host_mul_relin": (ci, [vp, vp, vp, vp, ci, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class DevBuf:
    """A uint64 buffer resident in HBM (freed with the engine's stream-ordered allocator)."""

    def __init__(self, eng, shape):
        self.eng, self.shape = eng, tuple(int(s) for s in shape)
        self.words = int(np.prod(self.shape))
        p = C.c_void_p()
        eng._ck(eng.lib.fl_dev_alloc(eng.h, self.words, C.byref(p)))
        self.ptr = p

    def upload(self, a):
        a = np.ascontiguousarray(a, np.uint64)
        assert a.size == self.words
        self.eng._ck(self.eng.lib.fl_dev_upload(self.eng.h, self.ptr, _ptr(a), self.words))
        self.eng.sync()
        return self

    def download(self):
        out = np.empty(self.shape, np.uint64)
        self.eng._ck(self.eng.lib.fl_dev_download(self.eng.h, _ptr(out), self.ptr, self.words))
        return out

    def free(self):
        if self.ptr is not None and self.eng.h is not None:
            self.eng.lib.fl_dev_free(self.eng.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """Raw-primitive front-end over the C-ABI (device-resident operands)."""

    def __init__(self, device=0, **params):
        self.lib = load_library()
        p = dict(ReferenceParams)
        p.update(params)
        self.params = p
        fp = fl_params(**p)
        h = C.c_void_p()
        self.h = None
        rc = self.lib.fl_ctx_create(C.byref(fp), device, C.byref(h))
        if rc:
            raise RuntimeError("fl_ctx_create: " + self.lib.fl_last_error().decode())
        self.h = h
        info = (C.c_int * 8)()
        self._ck(self.lib.fl_ctx_info(h, info))
        self.logN, self.L, self.K, self.alpha, self.dnum = (int(x) for x in info[:5])
        self.N = 1 << self.logN
        T = self.L + self.K
        self.moduli = np.zeros(T, np.uint64); self._ck(self.lib.fl_ctx_moduli(h, _ptr(self.moduli)))
        self.roots = np.zeros(T, np.uint64); self._ck(self.lib.fl_ctx_roots(h, _ptr(self.roots)))
        self.sf = np.zeros(self.L, np.float64); self._ck(self.lib.fl_ctx_scale_factors(h, _ptr(self.sf)))

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.lib.fl_last_error().decode())

    def close(self):
        if self.h is not None:
            self.lib.fl_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def sync(self): self._ck(self.lib.fl_sync(self.h))
    def stream(self): return self.lib.fl_ctx_stream(self.h)
    def buf(self, shape): return DevBuf(self, shape)
    def to_dev(self, a): return DevBuf(self, a.shape).upload(a)
    def galois(self, k): return int(self.lib.fl_galois_for_rotation(self.h, int(k)))
    def galois_conj(self): return int(self.lib.fl_galois_conj(self.h))

    @staticmethod
    def _midx(nl, midx):
        return np.arange(nl, dtype=np.int32) if midx is None else np.ascontiguousarray(midx, np.int32)

    # -- raw primitives (device buffers in, device buffers out)
    def ntt(self, d, midx=None):
        m = self._midx(d.shape[-2], midx); self._ck(self.lib.fl_raw_ntt(self.h, d.ptr, _ptr(m), len(m))); return d

    def intt(self, d, midx=None):
        m = self._midx(d.shape[-2], midx); self._ck(self.lib.fl_raw_intt(self.h, d.ptr, _ptr(m), len(m))); return d

    def ntt_batch(self, d, midx=None, inverse=False):
        """d: DevBuf [batch][limbs][N]; the whole batch in one launch pair."""
        m = self._midx(d.shape[-2], midx)
        self._ck(self.lib.fl_raw_ntt_batch(self.h, d.ptr, _ptr(m), len(m), d.shape[0], 1 if inverse else 0)); return d

    def _bin(self, fn, a, b, midx):
        m = self._midx(a.shape[-2], midx); out = self.buf(a.shape)
        self._ck(fn(self.h, out.ptr, a.ptr, b.ptr, _ptr(m), len(m))); return out

    def add(self, a, b, midx=None): return self._bin(self.lib.fl_raw_add, a, b, midx)
    def sub(self, a, b, midx=None): return self._bin(self.lib.fl_raw_sub, a, b, midx)
    def mul(self, a, b, midx=None): return self._bin(self.lib.fl_raw_mul, a, b, midx)

    def automorph(self, a, g):
        out = self.buf(a.shape); self._ck(self.lib.fl_raw_automorph(self.h, out.ptr, a.ptr, a.shape[-2], g)); return out

    def rescale(self, a):
        polys = a.shape[0] if len(a.shape) == 3 else 1
        l = a.shape[-2]
        out = self.buf(a.shape[:-2] + (l - 1, self.N))
        self._ck(self.lib.fl_raw_rescale(self.h, out.ptr, a.ptr, l, polys)); return out

    def modup(self, a, digit):
        l = a.shape[0]; out = self.buf((l + self.K, self.N))
        self._ck(self.lib.fl_raw_modup(self.h, out.ptr, a.ptr, l, digit)); return out

    def moddown(self, a):
        l = a.shape[0] - self.K; out = self.buf((l, self.N))
        self._ck(self.lib.fl_raw_moddown(self.h, out.ptr, a.ptr, l)); return out

    def keyswitch(self, a, evk):
        l = a.shape[0]; out = self.buf((2, l, self.N))
        self._ck(self.lib.fl_raw_keyswitch(self.h, out.ptr, a.ptr, evk.ptr, l)); return out

    def rotate(self, ct, g, evk, out=None):
        out = out or self.buf(ct.shape)
        self._ck(self.lib.fl_raw_rotate(self.h, out.ptr, ct.ptr, ct.shape[1], g, evk.ptr)); return out

    def rotate_batch(self, cts, g, evk, out=None):
        """cts: DevBuf [B][2][l][N]; one launch per key-switch stage for the whole batch."""
        out = out or self.buf(cts.shape)
        self._ck(self.lib.fl_raw_rotate_batch(self.h, out.ptr, cts.ptr, cts.shape[2], g, evk.ptr, cts.shape[0])); return out

    def host_rotate_batch(self, cts, g, evk, out=None, wait=True):
        """EvalRotate of host-resident ciphertexts [B][2][l][N]; wait=False returns at once (results valid after sync())."""
        out = np.empty_like(cts) if out is None else out
        fn = self.lib.fl_host_rotate_batch if wait else self.lib.fl_host_rotate_batch_async
        self._ck(fn(self.h, _ptr(out), _ptr(cts), cts.shape[2], g, evk.ptr, cts.shape[0])); return out

    def mul_relin(self, a, b, evk, out=None):
        out = out or self.buf(a.shape)
        self._ck(self.lib.fl_raw_mul_relin(self.h, out.ptr, a.ptr, b.ptr, a.shape[1], evk.ptr)); return out

    def mul_plain(self, ct, pt):
        out = self.buf(ct.shape)
        self._ck(self.lib.fl_raw_mul_plain(self.h, out.ptr, ct.ptr, pt.ptr, ct.shape[1])); return out

    # -- host-buffer entry points (H2D + kernels + D2H inside the call)
    def host_ntt(self, poly, inverse=False):
        a = np.ascontiguousarray(poly.copy(), np.uint64)
        self._ck(self.lib.fl_host_ntt(self.h, _ptr(a), a.shape[0], 1 if inverse else 0)); return a

    def host_rotate(self, ct, g, evk, out=None):
        out = np.empty_like(ct) if out is None else out
        self._ck(self.lib.fl_host_rotate(self.h, _ptr(out), _ptr(ct), ct.shape[1], g, evk.ptr)); return out

    def host_mul_relin(self, a, b, evk, out=None):
        out = np.empty_like(a) if out is None else out
        self._ck(self.lib.fl_host_mul_relin(self.h, _ptr(out), _ptr(a), _ptr(b), a.shape[1], evk.ptr)); return out
