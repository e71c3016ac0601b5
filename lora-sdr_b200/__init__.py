"""lora_sdr_b200 -- Python host binding over the C ABI (include/lorab200.h).

The product is ``liblorab200.so`` (hand-written sm_100a kernels + a plain C
ABI).  This module is a thin ctypes layer used by the tests, the bench and any
Python host; torch is only plumbing here (device memory, streams,
torch.distributed) and is optional for the host-buffer path.

There is no CPU fallback: if the library cannot be built/loaded or no B200 is
visible, construction of :class:`Demod` raises.

The directory is named ``lora-sdr_b200`` (task layout); import it through the
root-level shim module ``lora_sdr_b200``.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "liblorab200.so"
STAMP_PATH = PKG_DIR / "liblorab200.so.stamp"

TABLE_UP, TABLE_DOWN, TABLE_FINE, TABLE_TWIDDLE = 0, 1, 2, 3
FLAG_DEFAULT, FLAG_SIMPLE_ONLY, FLAG_INPUT_STABLE = 0, 1, 2
MOD_REFERENCE, MOD_IDEAL = 0, 1
FINE_STEPS = 128

RESULT_DTYPE = np.dtype(
    [("index", "<i4"), ("power", "<f4"), ("power_avg", "<f4"), ("f_index", "<f4")]
)

PREAMBLE_DTYPE = np.dtype(   # lorab200_preamble
    [("offset", "<u8"), ("aligned", "<u8"), ("phase", "<i4"), ("windows", "<i4"), ("snr_db", "<f4"), ("power_db", "<f4")]
)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # exact mode: FMA only where written explicitly
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-shared",
]

EXPORTS = [
    "lorab200_abi_version", "lorab200_strerror", "lorab200_last_error", "lorab200_create",
    "lorab200_destroy", "lorab200_sf", "lorab200_n", "lorab200_device", "lorab200_table_host",
    "lorab200_table_dev", "lorab200_set_table", "lorab200_tables_sync_from_dev",
    "lorab200_gen_chirp", "lorab200_demod_dev", "lorab200_demod_host", "lorab200_pack_symbols",
    "lorab200_kernel_launches", "lorab200_last_kernel", "lorab200_alloc_pinned",
    "lorab200_free_pinned", "lorab200_is_pinned", "lorab200_table_len", "lorab200_build_table", "lorab200_fine_advance",
    "lorab200_demod_multi_dev", "lorab200_mod_samples", "lorab200_mod_dev", "lorab200_decode_dev", "lorab200_rx_dev", "lorab200_rx_host", "lorab200_encode_host", "lorab200_decode_host", "lorab200_mod_host", "lorab200_encode_dev", "lorab200_encode_symbols",
    "lorab200_preamble_search_dev",
]


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp")) + sorted(CSRC.glob("*.cuh")) + \
        sorted(CSRC.glob("*.h")) + [REPO_ROOT / "include" / "lorab200.h"]


def _source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in _sources():
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def find_nvcc() -> str | None:
    return shutil.which("nvcc") or (
        "/usr/local/cuda/bin/nvcc" if Path("/usr/local/cuda/bin/nvcc").exists() else None)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/ into liblorab200.so for sm_100a (nvcc cross-compiles without a GPU).

    Rebuilds only when the sources (by content hash) changed, so the prebuilt
    library that travels to the GPU box is used as is.
    """
    want = _source_hash()
    if not force and LIB_PATH.exists() and STAMP_PATH.exists() and STAMP_PATH.read_text().strip() == want:
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        if LIB_PATH.exists():
            return LIB_PATH
        raise RuntimeError("nvcc not found and no prebuilt liblorab200.so")
    srcs = [str(p) for p in sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))]
    cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v"] + ["-o", str(LIB_PATH)] + srcs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    (PKG_DIR / "build.log").write_text(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout[-4000:])
    if verbose:
        print(r.stdout)
    STAMP_PATH.write_text(want)
    return LIB_PATH


_lib = None


def load_library() -> C.CDLL:
    """dlopen liblorab200.so and declare the C ABI. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("LORAB200_LIB", str(LIB_PATH)))   # LORAB200_LIB: kernel-tuning experiments
    if not path.exists():
        raise RuntimeError(f"{path} is not built; run __graft_entry__.build() (no CPU fallback exists)")
    L = C.CDLL(str(path))
    vp, sz, i32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_int32)
    L.lorab200_abi_version.restype = C.c_int
    L.lorab200_strerror.restype = C.c_char_p
    L.lorab200_strerror.argtypes = [C.c_int]
    L.lorab200_last_error.restype = C.c_char_p
    L.lorab200_last_error.argtypes = [vp]
    L.lorab200_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_uint32]
    L.lorab200_destroy.argtypes = [vp]
    L.lorab200_destroy.restype = None
    L.lorab200_sf.argtypes = [vp]
    L.lorab200_n.argtypes = [vp]
    L.lorab200_n.restype = sz
    L.lorab200_device.argtypes = [vp]
    L.lorab200_table_host.argtypes = [vp, C.c_int, C.POINTER(sz)]
    L.lorab200_table_host.restype = vp
    L.lorab200_table_dev.argtypes = [vp, C.c_int, C.POINTER(sz)]
    L.lorab200_table_dev.restype = vp
    L.lorab200_set_table.argtypes = [vp, C.c_int, vp]
    L.lorab200_tables_sync_from_dev.argtypes = [vp]
    L.lorab200_table_len.argtypes = [C.c_int, C.c_int]
    L.lorab200_table_len.restype = sz
    L.lorab200_build_table.argtypes = [C.c_int, C.c_int, vp]
    L.lorab200_fine_advance.argtypes = [C.c_int, C.c_int32, C.c_float, C.c_uint64]
    L.lorab200_fine_advance.restype = C.c_int32
    L.lorab200_gen_chirp.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                     C.POINTER(C.c_float)]
    L.lorab200_demod_dev.argtypes = [vp, vp, sz, sz, C.c_int, vp, vp, vp, vp, vp]
    L.lorab200_demod_host.argtypes = [vp, vp, sz, sz, C.c_int, vp, vp, vp, vp]
    L.lorab200_demod_multi_dev.argtypes = [vp, C.POINTER(vp), sz, sz, sz, C.c_int, vp, vp]
    L.lorab200_decode_dev.argtypes = [vp, vp, sz, sz, vp, vp, vp, sz, vp, vp]
    L.lorab200_rx_dev.argtypes = [vp, vp, sz, sz, sz, C.c_uint, C.c_float, sz, vp, vp, vp, vp, vp, vp, sz, vp]
    L.lorab200_rx_host.argtypes = [vp, vp, sz, sz, C.c_uint, C.c_float, sz, vp, vp, vp, vp, vp, vp, sz]
    L.lorab200_preamble_search_dev.argtypes = [vp, vp, sz, sz, C.c_float, C.c_int, vp, vp, C.c_uint32, vp]
    L.lorab200_encode_dev.argtypes = [vp, vp, sz, sz, vp, vp, vp, sz, vp, vp]
    L.lorab200_encode_symbols.restype = sz
    L.lorab200_encode_symbols.argtypes = [C.c_int, vp, sz]
    L.lorab200_mod_samples.restype = sz
    L.lorab200_mod_samples.argtypes = [C.c_int, C.c_int, sz, sz]
    L.lorab200_mod_dev.argtypes = [vp, vp, sz, sz, C.c_uint, sz, C.c_float, C.c_int, C.c_int, vp, vp]
    L.lorab200_pack_symbols.argtypes = [vp, sz, vp]
    L.lorab200_pack_symbols.restype = None
    L.lorab200_kernel_launches.argtypes = [vp]
    L.lorab200_kernel_launches.restype = C.c_ulonglong
    L.lorab200_last_kernel.argtypes = [vp]
    L.lorab200_last_kernel.restype = C.c_char_p
    L.lorab200_alloc_pinned.argtypes = [sz]
    L.lorab200_alloc_pinned.restype = vp
    L.lorab200_free_pinned.argtypes = [vp]
    L.lorab200_free_pinned.restype = None
    L.lorab200_is_pinned.argtypes = [vp]
    _lib = L
    return L


class Lorab200Error(RuntimeError):
    pass


class _Fine(C.Structure):
    _fields_ = [("index", C.c_int32), ("step", C.c_float)]


def gen_chirp(N: int, ovs: int, NN: int, f0: float, down: bool, ampl: float, phase: float):
    """ChirpGenerator.hpp:21-47 genChirp<float> (host). Returns (samples, phase_after)."""
    L = load_library()
    out = np.empty(NN, np.complex64)
    ph = C.c_float(phase)
    n = L.lorab200_gen_chirp(out.ctypes.data_as(C.c_void_p), N, ovs, NN, C.c_float(f0), int(down),
                             C.c_float(ampl), C.byref(ph))
    if n < 0:
        raise Lorab200Error(L.lorab200_strerror(n).decode())
    return out[:n], float(ph.value)


def fine_advance(sf: int, index: int, step: float, n_samples: int) -> int:
    """Closed-form advance of the fine-tune index recurrence (LoRaDemod.cpp:160-162)."""
    r = load_library().lorab200_fine_advance(sf, int(index), C.c_float(step), int(n_samples))
    if r < 0:
        raise Lorab200Error("bad arguments to lorab200_fine_advance")
    return int(r)


def build_table(sf: int, which: int) -> np.ndarray:
    """Host-only table recipe (no GPU needed): LoRaDemod.cpp:97-114 / kissfft.hh:17-22."""
    L = load_library()
    n = L.lorab200_table_len(sf, which)
    if n == 0:
        raise Lorab200Error("bad sf / table id")
    out = np.empty(n, np.complex64)
    st = L.lorab200_build_table(sf, which, out.ctypes.data_as(C.c_void_p))
    if st != 0:
        raise Lorab200Error(L.lorab200_strerror(st).decode())
    return out


class Demod:
    """One demodulator context (== one LoRaDemod block instance / one channel).

    Mirrors the hot-path part of ``LoRaDemod`` (LoRaDemod.cpp:68-117,157-172):
    the tables built in its constructor and the per-window dechirp + detect.
    """

    def __init__(self, sf: int, device: int = 0, flags: int = FLAG_DEFAULT):
        self.L = load_library()
        self.ctx = C.c_void_p()
        st = self.L.lorab200_create(C.byref(self.ctx), int(device), int(sf), int(flags))
        if st != 0:
            raise Lorab200Error(
                f"lorab200_create: {self.L.lorab200_strerror(st).decode()}: "
                f"{self.L.lorab200_last_error(None).decode()}")
        self.sf = sf
        self.n = 1 << sf
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.L.lorab200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int, what: str):
        if st != 0:
            raise Lorab200Error(f"{what}: {self.L.lorab200_strerror(st).decode()}: "
                                f"{self.L.lorab200_last_error(self.ctx).decode()}")

    # ---- tables -----------------------------------------------------------------
    def table(self, which: int) -> np.ndarray:
        n = C.c_size_t()
        p = self.L.lorab200_table_host(self.ctx, which, C.byref(n))
        buf = (C.c_float * (2 * n.value)).from_address(p)
        return np.frombuffer(buf, dtype=np.complex64).copy()

    def table_dev_ptr(self, which: int):
        n = C.c_size_t()
        p = self.L.lorab200_table_dev(self.ctx, which, C.byref(n))
        return p, n.value

    def set_table(self, which: int, data) -> None:
        data = np.ascontiguousarray(data, dtype=np.complex64)
        assert data.size == self.n
        self._check(self.L.lorab200_set_table(self.ctx, which, data.ctypes.data_as(C.c_void_p)), "set_table")

    def tables_sync_from_dev(self) -> None:
        self._check(self.L.lorab200_tables_sync_from_dev(self.ctx), "tables_sync_from_dev")

    @property
    def kernel_launches(self) -> int:
        return int(self.L.lorab200_kernel_launches(self.ctx))

    @property
    def last_kernel(self) -> str:
        return self.L.lorab200_last_kernel(self.ctx).decode()

    # ---- hot path, device-resident ----------------------------------------------
    def demod_dev(self, iq, n_windows: int, stride: int | None = None, table: int = TABLE_UP,
                  fine: tuple[int, float] | None = None, out=None, want_fft=False, want_dec=False,
                  stream=None):
        """iq: CUDA torch tensor (complex64, or float32 interleaved).  Asynchronous.

        Returns (results_tensor[int32 view, n_windows x 4 words], fine_index_end, fft, dec);
        use :func:`results_to_numpy` to get the structured records.
        """
        import torch

        assert iq.is_cuda and iq.is_contiguous()
        stride = self.n if stride is None else int(stride)
        dev = iq.device
        if out is None:
            out = torch.empty((n_windows, 4), dtype=torch.int32, device=dev)
        fft = torch.empty((n_windows, self.n), dtype=torch.complex64, device=dev) if want_fft else None
        dec = torch.empty((n_windows, self.n), dtype=torch.complex64, device=dev) if want_dec else None
        f = None
        if fine is not None:
            f = _Fine(int(fine[0]), float(fine[1]))
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        st = self.L.lorab200_demod_dev(
            self.ctx, C.c_void_p(iq.data_ptr()), n_windows, stride, table,
            C.byref(f) if f is not None else None, C.c_void_p(out.data_ptr()),
            C.c_void_p(fft.data_ptr()) if fft is not None else None,
            C.c_void_p(dec.data_ptr()) if dec is not None else None,
            C.c_void_p(stream))
        self._check(st, "demod_dev")
        return out, (f.index if f is not None else None), fft, dec

    def demod_multi_dev(self, segments, windows_per_segment: int, stride: int | None = None,
                        table: int = TABLE_UP, out=None, stream=None):
        """One launch over a list of CUDA tensors (one per channel), `windows_per_segment` windows each
        (lorab200_demod_multi_dev).  Returns the result tensor [len(segments) * windows_per_segment, 4]."""
        import torch

        assert len(segments) > 0 and all(t.is_cuda and t.is_contiguous() for t in segments)
        stride = self.n if stride is None else int(stride)
        dev = segments[0].device
        total = len(segments) * windows_per_segment
        if out is None:
            out = torch.empty((total, 4), dtype=torch.int32, device=dev)
        ptrs = (C.c_void_p * len(segments))(*[t.data_ptr() for t in segments])
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        st = self.L.lorab200_demod_multi_dev(self.ctx, ptrs, len(segments), windows_per_segment, stride, table,
                                             C.c_void_p(out.data_ptr()), C.c_void_p(stream))
        self._check(st, "demod_multi_dev")
        return out

    # ---- receiver: frame sync + symbol extraction on the device (LoRaDemod.cpp:145-327) -----------------
    def rx_dev(self, iq, n_samples=None, sync=0x12, thresh=-30.0, mtu=256, max_packets=16, state=None, stream=None):
        """iq: CUDA complex64 tensor [n_channels, samples].  Returns a dict of CUDA tensors: symbols
        [n_channels, max_packets, mtu] int16, info [n_channels, max_packets, 8] int32 words (length, error, power
        bits, snr bits, raw peak bits, 0, raw total as a double), n_packets [n_channels], consumed [n_channels], state (pass it back in to continue)."""
        import torch

        assert iq.is_cuda and iq.is_contiguous() and iq.dim() == 2
        nch, stride = iq.shape
        n_samples = stride if n_samples is None else int(n_samples)
        dev = iq.device
        if state is None:
            state = {"state": torch.zeros((nch, 14), dtype=torch.int32, device=dev),
                     "partial": torch.zeros((nch, mtu), dtype=torch.int16, device=dev)}
        symbols = torch.zeros((nch, max_packets, mtu), dtype=torch.int16, device=dev)
        info = torch.zeros((nch, max_packets, 8), dtype=torch.int32, device=dev)
        n_packets = torch.zeros(nch, dtype=torch.int32, device=dev)
        consumed = torch.zeros(nch, dtype=torch.int64, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        st = self.L.lorab200_rx_dev(self.ctx, C.c_void_p(iq.data_ptr()), nch, stride, n_samples, sync, thresh, mtu,
                                    C.c_void_p(state["state"].data_ptr()), C.c_void_p(state["partial"].data_ptr()),
                                    C.c_void_p(symbols.data_ptr()), C.c_void_p(info.data_ptr()),
                                    C.c_void_p(n_packets.data_ptr()), C.c_void_p(consumed.data_ptr()), max_packets,
                                    C.c_void_p(stream))
        self._check(st, "rx_dev")
        return dict(symbols=symbols, info=info, n_packets=n_packets, consumed=consumed, state=state)

    # ---- preamble search over all candidate offsets ---------------------------
    def preamble_search(self, iq, stride=None, thresh=-3.0, min_windows=8, max_found=256, stream=None):
        """iq: CUDA complex64 tensor [samples] (one stream).  Demodulates the windows at offsets 0, stride, ...
        in one launch and returns (records, n_found): records = numpy structured array (PREAMBLE_DTYPE) of the
        first min(n_found, max_found) preambles in stream order."""
        import torch

        assert iq.is_cuda and iq.is_contiguous() and iq.dim() == 1
        stride = (self.n // 4) if stride is None else int(stride)
        dev = iq.device
        found = torch.zeros((max(1, max_found), PREAMBLE_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        n_found = torch.zeros(1, dtype=torch.int32, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        st = self.L.lorab200_preamble_search_dev(self.ctx, C.c_void_p(iq.data_ptr()), iq.numel(), stride,
                                                 C.c_float(thresh), int(min_windows), C.c_void_p(found.data_ptr()),
                                                 C.c_void_p(n_found.data_ptr()), int(max_found), C.c_void_p(stream))
        self._check(st, "preamble_search_dev")
        torch.cuda.synchronize(dev)
        nf = int(n_found.item())
        rec = np.frombuffer(found.cpu().numpy().tobytes(), dtype=PREAMBLE_DTYPE)[: min(nf, max_found)].copy()
        return rec, nf

    def rx_host(self, iq_list, sync=0x12, thresh=-30.0, mtu=256, max_packets=16, state=None):
        """Host-buffer twin of :meth:`rx_dev`: iq_list = one contiguous complex64 numpy array per channel, all of
        the same length.  Returns a dict of numpy arrays (symbols, info, n_packets, consumed, state)."""
        nch = len(iq_list)
        arrs = [np.ascontiguousarray(a, dtype=np.complex64) for a in iq_list]
        n_samples = arrs[0].size
        assert all(a.size == n_samples for a in arrs)
        if state is None:
            state = {"state": np.zeros((nch, 14), np.int32), "partial": np.zeros((nch, mtu), np.int16)}
        ptrs = (C.c_void_p * nch)(*[a.ctypes.data for a in arrs])
        symbols = np.zeros((nch, max_packets, mtu), np.int16)
        info = np.zeros((nch, max_packets, 8), np.int32)
        n_packets = np.zeros(nch, np.int32)
        consumed = np.zeros(nch, np.uint64)
        vp = C.c_void_p
        st = self.L.lorab200_rx_host(self.ctx, ptrs, nch, n_samples, sync, thresh, mtu,
                                     vp(state["state"].ctypes.data), vp(state["partial"].ctypes.data),
                                     vp(symbols.ctypes.data), vp(info.ctypes.data), vp(n_packets.ctypes.data),
                                     vp(consumed.ctypes.data), max_packets)
        self._check(st, "rx_host")
        return dict(symbols=symbols, info=info, n_packets=n_packets, consumed=consumed, state=state)

    # ---- encoder (LoRaEncoder.cpp) ----------------------------------------------------------------
    def encode_dev(self, payloads, lengths=None, ppm=0, rdd=4, explicit_header=True, crc=True, whitening=True,
                   symbols_stride=None, stream=None):
        """payloads: CUDA uint8 tensor [n_packets, stride]; lengths: optional CUDA int32 [n_packets].
        Returns (symbols int16 [n_packets, symbols_stride], n_symbols int32 [n_packets])."""
        import torch

        assert payloads.is_cuda and payloads.is_contiguous() and payloads.dim() == 2 and payloads.dtype == torch.uint8
        npk, stride = payloads.shape
        cfg = (C.c_int32 * 5)(ppm, rdd, int(explicit_header), int(crc), int(whitening))
        if symbols_stride is None:
            symbols_stride = int(self.L.lorab200_encode_symbols(self.sf, C.cast(cfg, C.c_void_p), stride))
        out = torch.zeros((npk, symbols_stride), dtype=torch.int16, device=payloads.device)
        nsy = torch.empty(npk, dtype=torch.int32, device=payloads.device)
        if stream is None:
            stream = torch.cuda.current_stream(payloads.device).cuda_stream
        st = self.L.lorab200_encode_dev(self.ctx, C.c_void_p(payloads.data_ptr()), npk, stride,
                                        C.c_void_p(lengths.data_ptr()) if lengths is not None else None,
                                        C.cast(cfg, C.c_void_p), C.c_void_p(out.data_ptr()), symbols_stride,
                                        C.c_void_p(nsy.data_ptr()), C.c_void_p(stream))
        self._check(st, "encode_dev")
        return out, nsy

    # ---- decoder (LoRaDecoder.cpp) ----------------------------------------------------------------
    def decode_dev(self, symbols, n_symbols=None, ppm=0, rdd=4, explicit_header=True, hdr=False, data_length=8,
                   crcc=False, error_check=False, bytes_stride=264, stream=None):
        """symbols: CUDA int16/uint16 tensor [n_packets, stride]; n_symbols: optional CUDA int32 [n_packets].
        Returns (bytes uint8 [n_packets, bytes_stride], lengths int32 [n_packets])."""
        import torch

        assert symbols.is_cuda and symbols.is_contiguous() and symbols.dim() == 2 and symbols.element_size() == 2
        npk, stride = symbols.shape
        cfg = (C.c_int32 * 7)(ppm, rdd, int(explicit_header), int(hdr), data_length, int(crcc), int(error_check))
        out = torch.zeros((npk, bytes_stride), dtype=torch.uint8, device=symbols.device)
        lengths = torch.empty(npk, dtype=torch.int32, device=symbols.device)
        if stream is None:
            stream = torch.cuda.current_stream(symbols.device).cuda_stream
        st = self.L.lorab200_decode_dev(self.ctx, C.c_void_p(symbols.data_ptr()), npk, stride,
                                        C.c_void_p(n_symbols.data_ptr()) if n_symbols is not None else None,
                                        C.cast(cfg, C.c_void_p), C.c_void_p(out.data_ptr()), bytes_stride,
                                        C.c_void_p(lengths.data_ptr()), C.c_void_p(stream))
        self._check(st, "decode_dev")
        return out, lengths

    # ---- modulator (LoRaMod.cpp) ------------------------------------------------------------------
    def mod_dev(self, symbols, sync: int = 0x12, padding: int = 1, ampl: float = 0.3, ovs: int = 1,
                mode: int = 0, out=None, stream=None):
        """symbols: CUDA int16/uint16 tensor [n_packets, syms_per_packet] -> complex64 [n_packets, samples]."""
        import torch

        assert symbols.is_cuda and symbols.is_contiguous() and symbols.dim() == 2 and symbols.element_size() == 2
        npk, nsy = symbols.shape
        spp = self.L.lorab200_mod_samples(self.sf, ovs, nsy, padding)
        if out is None:
            out = torch.empty((npk, spp), dtype=torch.complex64, device=symbols.device)
        if stream is None:
            stream = torch.cuda.current_stream(symbols.device).cuda_stream
        st = self.L.lorab200_mod_dev(self.ctx, C.c_void_p(symbols.data_ptr()), npk, nsy, sync, padding, ampl, ovs, mode,
                                     C.c_void_p(out.data_ptr()), C.c_void_p(stream))
        self._check(st, "mod_dev")
        return out

    # ---- hot path, host buffers (what a Pothos work() calls) -----------------------
    def demod_host(self, iq, n_windows: int, stride: int | None = None, table: int = TABLE_UP,
                   fine: tuple[int, float] | None = None, want_fft=False, want_dec=False, out=None):
        """iq: numpy complex64 (any shape, contiguous) or a CPU torch tensor (pinned is faster)."""
        stride = self.n if stride is None else int(stride)
        ptr, nsamp = _host_ptr(iq)
        if n_windows:
            assert nsamp >= (n_windows - 1) * stride + self.n, "iq too short"
        if out is None:
            out = np.empty(n_windows, RESULT_DTYPE)
        fft = np.empty((n_windows, self.n), np.complex64) if want_fft else None
        dec = np.empty((n_windows, self.n), np.complex64) if want_dec else None
        f = _Fine(int(fine[0]), float(fine[1])) if fine is not None else None
        st = self.L.lorab200_demod_host(
            self.ctx, ptr, n_windows, stride, table, C.byref(f) if f is not None else None,
            _host_ptr(out)[0],
            fft.ctypes.data_as(C.c_void_p) if fft is not None else None,
            dec.ctypes.data_as(C.c_void_p) if dec is not None else None)
        self._check(st, "demod_host")
        return out, (f.index if f is not None else None), fft, dec

    def pack_symbols(self, results: np.ndarray) -> np.ndarray:
        sym = np.empty(results.shape[0], np.int16)
        self.L.lorab200_pack_symbols(results.ctypes.data_as(C.c_void_p), results.shape[0],
                                     sym.ctypes.data_as(C.c_void_p))
        return sym


def _host_ptr(a):
    """(void*, n_complex) of a numpy array or CPU torch tensor."""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        n = a.size if a.dtype == np.complex64 else a.nbytes // 8
        return a.ctypes.data_as(C.c_void_p), n
    # torch CPU tensor
    assert not a.is_cuda and a.is_contiguous()
    return C.c_void_p(a.data_ptr()), a.numel() * a.element_size() // 8


class PinnedArray:
    """complex64 numpy array in page-locked host memory from lorab200_alloc_pinned (cudaMallocHost):
    the buffers a host should hand to demod_host -- copies then run at full PCIe rate
    (measured here: ~52 GB/s against ~27 GB/s for torch's pin_memory())."""

    def __init__(self, shape):
        L = load_library()
        self.shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = int(np.prod(self.shape))
        self._nbytes = n * 8
        self._ptr = L.lorab200_alloc_pinned(self._nbytes)
        if not self._ptr:
            raise Lorab200Error("lorab200_alloc_pinned failed")
        buf = (C.c_float * (2 * n)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=np.complex64).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            load_library().lorab200_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def results_to_numpy(out_tensor) -> np.ndarray:
    """Device results tensor (n x 4 int32 words) -> structured numpy records."""
    return out_tensor.cpu().numpy().view(RESULT_DTYPE).reshape(-1)
