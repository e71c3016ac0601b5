"""ctypes loader for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module (as the checker or the timed
CPU baseline).  The product package never does.

Two libraries:
  * ``liblora_oracle.so``  -- our C restatement (oracle/lora_oracle.c), built by
    ``make -C oracle`` anywhere gcc exists.
  * ``_ref/libloraref.so`` -- the real reference headers compiled in place from
    /root/reference (oracle/ref_shim.cpp).  Built only where /root/reference
    exists; the prebuilt file travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_ORACLE = HERE / "liblora_oracle.so"
LIB_REF = HERE / "_ref" / "libloraref.so"
LIB_REF_O3 = HERE / "_ref" / "libloraref_o3.so"   # -O3 -march=x86-64-v3 courtesy build (same results)

RESULT_DTYPE = np.dtype(
    [("index", "<i4"), ("power", "<f4"), ("power_avg", "<f4"), ("f_index", "<f4")]
)


def build(force: bool = False) -> None:
    """Compile the restatement, the host build of the product's codec / state-machine headers and
    (when /root/reference exists) the real reference + loopback chain.  Always delegates to make,
    which knows every dependency (block shell, ministream headers, product headers); an
    up-to-date tree costs a few milliseconds."""
    args = ["make", "-C", str(HERE)] + (["-B"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


class _Plan(C.Structure):
    _fields_ = [
        ("sf", C.c_int),
        ("n", C.c_int),
        ("nstages", C.c_int),
        ("radix", C.c_int * 16),
        ("remainder", C.c_int * 16),
        ("twiddles", C.c_void_p),
        ("power_scale", C.c_float),
        ("rev", C.c_void_p),
    ]


def _c64(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.complex64)
    return a


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Our C restatement (kind == "port")."""

    kind = "port"

    def __init__(self):
        build()
        self.lib = C.CDLL(str(LIB_ORACLE))
        L = self.lib
        L.lo_plan_create.restype = C.POINTER(_Plan)
        L.lo_plan_create.argtypes = [C.c_int]
        L.lo_plan_destroy.argtypes = [C.POINTER(_Plan)]
        L.lo_twiddles.argtypes = [C.c_int, C.c_void_p]
        L.lo_tables_demod.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.lo_table_fine.argtypes = [C.c_int, C.c_void_p]
        L.lo_gen_chirp.restype = C.c_int
        L.lo_gen_chirp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_int, C.c_float, C.POINTER(C.c_float)]
        L.lo_fft.argtypes = [C.POINTER(_Plan), C.c_void_p, C.c_void_p]
        L.lo_detect.restype = C.c_size_t
        L.lo_detect.argtypes = [C.POINTER(_Plan), C.c_void_p, C.POINTER(C.c_float),
                                C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
        L.lo_fine_next.restype = C.c_int32
        L.lo_fine_next.argtypes = [C.c_int32, C.c_float, C.c_int32]
        L.lo_demod_windows.argtypes = [C.POINTER(_Plan), C.c_void_p, C.c_size_t, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self._plans = {}

    def plan(self, sf: int):
        if sf not in self._plans:
            self._plans[sf] = self.lib.lo_plan_create(sf)
        return self._plans[sf]

    def plan_info(self, sf: int):
        p = self.plan(sf).contents
        return list(p.radix[: p.nstages]), list(p.remainder[: p.nstages]), float(p.power_scale)

    def twiddles(self, sf: int) -> np.ndarray:
        n = 1 << sf
        tw = np.empty(n, np.complex64)
        self.lib.lo_twiddles(n, _ptr(tw))
        return tw

    def tables_demod(self, sf: int):
        n = 1 << sf
        up, down = np.empty(n, np.complex64), np.empty(n, np.complex64)
        self.lib.lo_tables_demod(sf, _ptr(up), _ptr(down))
        return up, down

    def table_fine(self, sf: int) -> np.ndarray:
        fine = np.empty((1 << sf) * 128, np.complex64)
        self.lib.lo_table_fine(sf, _ptr(fine))
        return fine

    def gen_chirp(self, N, ovs, NN, f0, down, ampl, phase):
        out = np.empty(NN, np.complex64)
        ph = C.c_float(phase)
        n = self.lib.lo_gen_chirp(_ptr(out), N, ovs, NN, C.c_float(f0), int(down),
                                  C.c_float(ampl), C.byref(ph))
        return out[:n], float(ph.value)

    def fft(self, sf: int, x) -> np.ndarray:
        x = _c64(x)
        out = np.empty_like(x)
        self.lib.lo_fft(self.plan(sf), _ptr(x), _ptr(out))
        return out

    def detect(self, sf: int, x, want_fft=False):
        x = _c64(x)
        p, pa, fi = C.c_float(), C.c_float(), C.c_float()
        fo = np.empty_like(x) if want_fft else None
        idx = self.lib.lo_detect(self.plan(sf), _ptr(x), C.byref(p), C.byref(pa), C.byref(fi), _ptr(fo))
        r = (int(idx), float(p.value), float(pa.value), float(fi.value))
        return r + (fo,) if want_fft else r

    def fine_next(self, idx, step, length):
        return int(self.lib.lo_fine_next(int(idx), C.c_float(step), int(length)))

    def demod_windows(self, sf, iq, n_windows, chirp, stride=None, fine=None, fine_idx=0,
                      fine_step=0.0, want_fft=False, want_dec=False, nthreads=1):
        """Returns (results[n_windows] structured, fine_idx_end, fft|None, dec|None)."""
        n = 1 << sf
        iq = _c64(iq).reshape(-1)
        stride = n if stride is None else int(stride)
        if n_windows:
            assert iq.size >= (n_windows - 1) * stride + n
        chirp = _c64(chirp)
        fine = None if fine is None else _c64(fine)
        out = np.zeros(n_windows, RESULT_DTYPE)
        fo = np.empty((n_windows, n), np.complex64) if want_fft else None
        do = np.empty((n_windows, n), np.complex64) if want_dec else None
        idx = C.c_int32(int(fine_idx))
        self._call_demod(sf, iq, n_windows, stride, chirp, fine, idx, fine_step, out, fo, do, nthreads)
        return out, int(idx.value), fo, do

    def _call_demod(self, sf, iq, nw, stride, chirp, fine, idx, fine_step, out, fo, do, nthreads):
        self.lib.lo_demod_windows(self.plan(sf), _ptr(iq), nw, stride, _ptr(chirp), _ptr(fine),
                                  C.byref(idx), C.c_float(fine_step), _ptr(out), _ptr(fo), _ptr(do),
                                  int(nthreads))


class Reference(Oracle):
    """The real reference headers compiled in place (kind == "reference").

    Table recipes (``tables_demod``/``table_fine``/``twiddles``) come from the
    restatement unless the shim was built with LoRaDemod.cpp (``has_demod_tables``).
    NOTE: fine_step here is ``_finefreqError`` itself; the shim multiplies by
    _fineSteps (128) with the reference's own operand types.
    """

    kind = "reference"

    def __init__(self, o3: bool = False):
        super().__init__()
        lib = LIB_REF_O3 if o3 else LIB_REF
        if not lib.exists():
            raise FileNotFoundError(f"{lib} not built (needs /root/reference)")
        self.ref = C.CDLL(str(lib))
        R = self.ref
        R.ref_fft.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        R.ref_detect.restype = C.c_size_t
        R.ref_detect.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), C.c_void_p]
        R.ref_gen_chirp.restype = C.c_int
        R.ref_gen_chirp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                    C.c_float, C.POINTER(C.c_float)]
        R.ref_demod_windows.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                        C.c_void_p, C.POINTER(C.c_int32), C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int]
        R.ref_demod_tables.restype = C.c_int
        R.ref_demod_tables.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

    @property
    def has_demod_tables(self) -> bool:
        return bool(self.ref.ref_demod_tables(7, None, None, None))

    def ref_tables(self, sf):
        n = 1 << sf
        up, down = np.empty(n, np.complex64), np.empty(n, np.complex64)
        fine = np.empty(n * 128, np.complex64)
        ok = self.ref.ref_demod_tables(sf, _ptr(up), _ptr(down), _ptr(fine))
        return (up, down, fine) if ok else None

    def gen_chirp(self, N, ovs, NN, f0, down, ampl, phase):
        out = np.empty(NN, np.complex64)
        ph = C.c_float(phase)
        n = self.ref.ref_gen_chirp(_ptr(out), N, ovs, NN, C.c_float(f0), int(down),
                                   C.c_float(ampl), C.byref(ph))
        return out[:n], float(ph.value)

    def fft(self, sf, x):
        x = _c64(x)
        out = np.empty_like(x)
        self.ref.ref_fft(sf, _ptr(x), _ptr(out))
        return out

    def detect(self, sf, x, want_fft=False):
        x = _c64(x)
        p, pa, fi = C.c_float(), C.c_float(), C.c_float()
        fo = np.empty_like(x) if want_fft else None
        idx = self.ref.ref_detect(sf, _ptr(x), C.byref(p), C.byref(pa), C.byref(fi), _ptr(fo))
        r = (int(idx), float(p.value), float(pa.value), float(fi.value))
        return r + (fo,) if want_fft else r

    def _call_demod(self, sf, iq, nw, stride, chirp, fine, idx, fine_step, out, fo, do, nthreads):
        # the shim takes _finefreqError, not the pre-multiplied step
        self.ref.ref_demod_windows(sf, _ptr(iq), nw, stride, _ptr(chirp), _ptr(fine), C.byref(idx),
                                   C.c_float(np.float32(fine_step) / np.float32(128.0)), _ptr(out),
                                   _ptr(fo), _ptr(do), int(nthreads))


def have_reference() -> bool:
    build()
    return LIB_REF.exists()
