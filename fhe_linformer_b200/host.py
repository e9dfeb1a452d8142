"""ctypes binding of host/flhost.h: the FHEController veneer and the encrypted Linformer forward (libflhost.so).

This is the call a user of the reference makes (FHEController + main.cpp's pipeline), re-backed by the CUDA engine.
No fallback: loading fails loudly when the libraries have not been built."""
import ctypes as C
import os

import numpy as np

from . import capi
from .ckks import CKKS, Elem

HOST_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libflhost.so")
CHECKPOINT_FN = C.CFUNCTYPE(None, C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)

# every rotation index the circuit uses
CIRCUIT_ROTATIONS = [1 << i for i in range(14)] + [-(1 << i) for i in range(7)] + [-512]

_hlib = None


def load_host_library():
    global _hlib
    if _hlib is not None:
        return _hlib
    capi.load_library()
    if not os.path.exists(HOST_LIB_PATH):
        raise RuntimeError(f"{HOST_LIB_PATH} not found: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(HOST_LIB_PATH)
    vp, ci = C.c_void_p, C.c_int
    L.flh_last_error.restype = C.c_char_p
    L.flh_new.restype = vp; L.flh_new.argtypes = [ci, C.c_ulonglong]
    L.flh_free.argtypes = [vp]
    L.flh_native.restype = vp; L.flh_native.argtypes = [vp]
    L.flh_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.flh_rotation_key_bytes.restype = C.c_double; L.flh_rotation_key_bytes.argtypes = [vp]
    L.flh_generate.argtypes = [vp, ci, vp, ci, ci, ci]
    L.flh_load.argtypes = [vp, C.c_char_p, ci]
    L.flh_info.argtypes = [vp, C.POINTER(ci), C.POINTER(ci)]
    L.flh_forward_many.argtypes = [vp, C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), ci, ci, ci, ci, vp, C.POINTER(ci)]
    L.flh_forward.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, ci, ci, ci, vp, CHECKPOINT_FN, vp, C.c_char_p, ci, vp, C.POINTER(ci), C.POINTER(ci)]
    L.flh_invoke.argtypes = [vp, C.c_char_p, vp, ci, vp, ci, vp, ci, vp, ci, vp, ci, C.POINTER(ci)]
    _hlib = L
    return L


class _Borrowed(CKKS):
    """CKKS front-end over a context owned by an FHEController (never destroys it)."""

    def __init__(self, handle):
        self.lib = capi.load_library()
        from .ckks import _SIGS
        for name, (res, args) in _SIGS.items():
            f = getattr(self.lib, name); f.restype, f.argtypes = res, args
        self.h = C.c_void_p(handle)
        info = (C.c_int * 8)()
        self._ck(self.lib.fl_ctx_info(self.h, info))
        self.logN, self.L, self.K, self.alpha, self.dnum = (int(x) for x in info[:5])
        self.N = 1 << self.logN
        T = self.L + self.K
        self.moduli = np.zeros(T, np.uint64); self._ck(self.lib.fl_ctx_moduli(self.h, capi._ptr(self.moduli)))
        self.sf = np.zeros(self.L, np.float64); self._ck(self.lib.fl_ctx_scale_factors(self.h, capi._ptr(self.sf)))

    def close(self):
        self.h = None


class FHEController:
    """The reference's FHEController (src/FHEController.h:22-162) on the B200 engine."""

    def __init__(self, device=0, key_seed=0, root=None, **options):
        """options: cache_gb, auto_rotation_keys, batch_rows, hoist_ladders, max_rows_per_batch (host/FHEController.h)."""
        self.hl = load_host_library()
        if root is not None:
            os.environ["FHE_LINFORMER_ROOT"] = root
        self.root = root
        self.h = C.c_void_p(self.hl.flh_new(device, key_seed))
        self.ckks = None
        for k, v in options.items():
            self.set_option(k, v)

    def set_option(self, name, value):
        self._ck(self.hl.flh_set_option(self.h, name.encode(), float(value)))

    def rotation_key_bytes(self):
        return float(self.hl.flh_rotation_key_bytes(self.h))

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.hl.flh_last_error().decode())

    def generate(self, log_ring=0, rotations=CIRCUIT_ROTATIONS, bootstrap_slots=16384, serialize=False):
        """generate_context + generate_bootstrapping_and_rotation_keys (main.cpp:82-85)."""
        r = np.ascontiguousarray(rotations, np.int32)
        self._ck(self.hl.flh_generate(self.h, log_ring, capi._ptr(r), len(r), bootstrap_slots, 1 if serialize else 0))
        self.ckks = _Borrowed(self.hl.flh_native(self.h))
        return self

    def load(self, rotation_file="rotation_keys.txt", bootstrap_slots=16384):
        """load_context + load_bootstrapping_and_rotation_keys (main.cpp:88-89)."""
        self._ck(self.hl.flh_load(self.h, rotation_file.encode(), bootstrap_slots))
        self.ckks = _Borrowed(self.hl.flh_native(self.h))
        return self

    @property
    def circuit_depth(self):
        d, n = C.c_int(), C.c_int(); self.hl.flh_info(self.h, C.byref(d), C.byref(n)); return d.value

    @property
    def num_slots(self):
        d, n = C.c_int(), C.c_int(); self.hl.flh_info(self.h, C.byref(d), C.byref(n)); return n.value

    def forward(self, dirs, token_limit=0, dead_work=True, classes=20, checkpoints=None, encrypted_projection=False, all_tokens=False, packed=False):
        """encoder1 -> pooler -> classifier -> decrypt (main.cpp:105-123) on the text files under `dirs`
        ({"weights", "input", "tokens"}).  Returns (logits, {stage: seconds}, S)."""
        logits = np.zeros(classes)
        names = C.create_string_buffer(4096); secs = np.zeros(32); nt = C.c_int(32); toks = C.c_int(0)

        def sink(name, ptr, n, level, _user):
            if checkpoints is not None:
                checkpoints[name.decode()] = (np.ctypeslib.as_array(ptr, shape=(n,)).copy(), level)
        cb = CHECKPOINT_FN(sink) if checkpoints is not None else C.cast(None, CHECKPOINT_FN)
        self._ck(self.hl.flh_forward(self.h, dirs["weights"].encode(), dirs["input"].encode(), dirs["tokens"].encode(), token_limit,
                                     (1 if dead_work else 0) | (2 if encrypted_projection else 0) | (4 if all_tokens else 0) | (8 if packed else 0), classes, capi._ptr(logits), cb, None, names, len(names), capi._ptr(secs),
                                     C.byref(nt), C.byref(toks)))
        stage = dict(zip(names.value.decode().split("\n"), secs[:nt.value].tolist()))
        return logits, stage, toks.value

    def forward_many(self, dirs_list, token_limit=0, dead_work=False, classes=20, packed=True):
        """The forward on several samples in one pass (flh_forward_many): every ciphertext carries one element per sample.
        `dirs_list`: one {"weights", "input", "tokens"} per sample, same weights folder and same number of rows.  Returns
        (logits[samples][classes], S)."""
        n = len(dirs_list)
        logits = np.zeros((n, classes)); toks = C.c_int(0)
        ins = (C.c_char_p * n)(*[d["input"].encode() for d in dirs_list]); tks = (C.c_char_p * n)(*[d["tokens"].encode() for d in dirs_list])
        self._ck(self.hl.flh_forward_many(self.h, dirs_list[0]["weights"].encode(), ins, tks, n, token_limit,
                                          (1 if dead_work else 0) | (8 if packed else 0), classes, capi._ptr(logits), C.byref(toks)))
        return logits, toks.value

    def invoke(self, method, cts=(), pts=(), ints=(), reals=(), out_cap=1024):
        """Call an FHEController method by name on C-ABI handles (layout tests)."""
        ca = (C.c_void_p * max(1, len(cts)))(*[e.h for e in cts])
        pa = (C.c_void_p * max(1, len(pts)))(*[(e.h if e is not None else None) for e in pts])
        ia = np.ascontiguousarray(ints, np.int32); ra = np.ascontiguousarray(reals, np.float64)
        out = (C.c_void_p * out_cap)(); n = C.c_int(0)
        self._ck(self.hl.flh_invoke(self.h, method.encode(), ca, len(cts), pa, len(pts), capi._ptr(ia), len(ia), capi._ptr(ra), len(ra), out, out_cap,
                                    C.byref(n)))
        return [Elem(self.ckks, C.c_void_p(out[i])) for i in range(n.value)]

    def close(self):
        if self.h is not None:
            if self.ckks is not None:
                self.ckks.close()
            self.hl.flh_free(self.h)
            self.h = None
