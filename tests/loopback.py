"""ctypes access to oracle/_ref/libloraloop.so (the reference chain + our demod block shell)."""
import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libloraloop.so"


class LoopCfg(C.Structure):
    _fields_ = [("sf", C.c_int), ("cr", C.c_int), ("use_b200", C.c_int), ("debug_ports", C.c_int),
                ("padding", C.c_int), ("mtu", C.c_int), ("sync", C.c_int), ("ampl", C.c_float),
                ("sigma", C.c_float), ("thresh", C.c_float), ("seed", C.c_ulonglong)]


class LoopOut(C.Structure):
    _fields_ = [("n_packets", C.c_int), ("lengths", C.c_int * 64), ("bytes", C.c_ubyte * (64 * 256)),
                ("n_sym_packets", C.c_int), ("sym_lengths", C.c_int * 64), ("symbols", C.c_short * (64 * 1024)),
                ("dropped", C.c_ulonglong), ("launches", C.c_ulonglong), ("samples", C.c_ulonglong),
                ("n_labels", C.c_int), ("labels", (C.c_char * 24) * 256), ("label_index", C.c_ulonglong * 256),
                ("n_signals", C.c_int), ("signal_power", C.c_float * 64), ("signal_error", C.c_int * 64),
                ("error", C.c_char * 256), ("calls", C.c_ulonglong),
                ("seconds", C.c_double), ("n_decoded", C.c_ulonglong), ("windows", C.c_ulonglong),
                ("pinned_inputs", C.c_ulonglong)]


def available():
    return LIB.exists()


def _lib():
    L = C.CDLL(str(LIB))
    L.loopback_tx.argtypes = [C.POINTER(LoopCfg), C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_void_p,
                              C.c_ulonglong, C.POINTER(C.c_ulonglong), C.c_char_p]
    L.loopback_rx.argtypes = [C.POINTER(LoopCfg), C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.POINTER(LoopOut)]
    L.loopback_encode.argtypes = [C.POINTER(LoopCfg), C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                  C.c_char_p]
    L.loopback_encode_cfg.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_char_p]
    L.loopback_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_char_p]
    L.loopback_mod.argtypes = [C.POINTER(LoopCfg), C.c_void_p, C.c_int, C.c_void_p, C.c_ulonglong,
                               C.POINTER(C.c_ulonglong), C.c_char_p]
    return L


def use_b200_codec(mask):
    """Which encoder (bit 0) / modulator (bit 1) / decoder (bit 2) the helpers below instantiate:
    our block shells (lora-sdr_b200/host/blocks/LoRaCodecB200.cpp) instead of the reference blocks."""
    L = _lib()
    L.loopback_use_b200_codec.argtypes = [C.c_int]
    L.loopback_use_b200_codec(int(mask))


def encode(sf, cr, payload):
    """payload bytes -> symbols (uint16) of one packet, by the reference LoRaEncoder."""
    L = _lib()
    cfg = LoopCfg(sf, cr, 0, 0, 0, 0, 0x12, 1.0, 0.0, 0.0, 0)
    syms = np.empty(4096, np.uint16)
    n = C.c_int()
    err = C.create_string_buffer(256)
    if L.loopback_encode(C.byref(cfg), payload, len(payload), syms.ctypes.data_as(C.c_void_p), syms.size, C.byref(n), err):
        raise RuntimeError(err.value.decode())
    return syms[: n.value].copy()


def encode_cfg(cfg6, payload):
    """payload -> symbols by the reference LoRaEncoder; cfg6 = (sf, ppm, rdd, explicit, crc, whitening)."""
    L = _lib()
    cfg = (C.c_int * 6)(*[int(v) for v in cfg6])
    syms = np.empty(4096, np.uint16)
    n = C.c_int()
    err = C.create_string_buffer(256)
    if L.loopback_encode_cfg(cfg, payload, len(payload), syms.ctypes.data_as(C.c_void_p), syms.size, C.byref(n), err):
        raise RuntimeError(err.value.decode())
    return syms[: n.value].copy()


def decode(cfg8, syms):
    """one packet of symbols -> (length, bytes) by the reference LoRaDecoder; cfg8 = (sf, ppm, rdd, explicit, hdr,
    data_length, crcc, error_check).  length -1 = dropped, -2 = nothing produced."""
    L = _lib()
    cfg = (C.c_int * 8)(*[int(v) for v in cfg8])
    syms = np.ascontiguousarray(syms, np.uint16)
    out = np.zeros(8192, np.uint8)
    n = C.c_int()
    err = C.create_string_buffer(256)
    if L.loopback_decode(cfg, syms.ctypes.data_as(C.c_void_p), syms.size, out.ctypes.data_as(C.c_void_p), out.size,
                         C.byref(n), err):
        raise RuntimeError(err.value.decode())
    return n.value, bytes(out[: max(n.value, 0)])


def modulate(sf, syms, padding=1, sync=0x12, ampl=1.0, ovs=1):
    """one packet of symbols -> IQ (complex64), by the reference LoRaMod."""
    L = _lib()
    cfg = LoopCfg(sf, 0, 0, 0, padding, ovs, sync, ampl, 0.0, 0.0, 0)
    syms = np.ascontiguousarray(syms, np.uint16)
    cap = (16 + syms.size + padding) * (1 << sf) * ovs
    iq = np.empty(cap, np.complex64)
    n = C.c_ulonglong()
    err = C.create_string_buffer(256)
    if L.loopback_mod(C.byref(cfg), syms.ctypes.data_as(C.c_void_p), syms.size, iq.ctypes.data_as(C.c_void_p), cap,
                      C.byref(n), err):
        raise RuntimeError(err.value.decode())
    return iq[: n.value].copy()


def transmit(sf, cr, packets, padding=512, sync=0x12, ampl=1.0, sigma=0.0, seed=1):
    """packets -> reference encoder -> reference modulator -> + N(0, sigma^2) noise.  Returns complex64 samples."""
    L = _lib()
    cfg = LoopCfg(sf, cr, 0, 0, padding, 0, sync, ampl, sigma, 0.0, seed)
    blob = b"".join(packets)
    lens = (C.c_int * len(packets))(*[len(p) for p in packets])
    n_sym_max = sum(8 + 2 * (len(p) + 4) * 8 // max(1, sf - 2) + 16 for p in packets)
    cap = (n_sym_max + len(packets) * (padding + 16)) * (1 << sf)
    iq = np.empty(cap, np.complex64)
    n = C.c_ulonglong()
    err = C.create_string_buffer(256)
    rc = L.loopback_tx(C.byref(cfg), blob, lens, len(packets), iq.ctypes.data_as(C.c_void_p), cap, C.byref(n), err)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return iq[: n.value].copy()


def receive(sf, cr, iq, use_b200=False, debug_ports=True, mtu=512, sync=0x12, thresh=-30.0, feed=0):
    """IQ -> demod (reference LoRaDemod.cpp or our block shell) -> reference decoder."""
    L = _lib()
    cfg = LoopCfg(sf, cr, int(use_b200), int(debug_ports), 0, mtu, sync, 1.0, 0.0, thresh, 0)
    iq = np.ascontiguousarray(iq, np.complex64)
    out = LoopOut()
    rc = L.loopback_rx(C.byref(cfg), iq.ctypes.data_as(C.c_void_p), iq.size, feed, C.byref(out))
    if rc != 0:
        raise RuntimeError(out.error.decode())
    dec, o = [], 0
    raw = bytes(out.bytes)
    for i in range(out.n_packets):
        dec.append(raw[o:o + out.lengths[i]])
        o += out.lengths[i]
    syms, o = [], 0
    sarr = np.ctypeslib.as_array(out.symbols)
    for i in range(out.n_sym_packets):
        syms.append(sarr[o:o + out.sym_lengths[i]].copy())
        o += out.sym_lengths[i]
    labels = [(out.labels[i].value.decode(), int(out.label_index[i])) for i in range(out.n_labels)]
    return dict(decoded=dec, symbols=syms, dropped=int(out.dropped), launches=int(out.launches), calls=int(out.calls), seconds=float(out.seconds), n_decoded=int(out.n_decoded), windows=int(out.windows),
                pinned_inputs=int(out.pinned_inputs),
                samples=int(out.samples), labels=labels,
                power=[float(out.signal_power[i]) for i in range(out.n_signals)],
                error=[int(out.signal_error[i]) for i in range(out.n_signals)])


def receive_multi(sf, streams, mtu=512, sync=0x12, thresh=-30.0, feed=0):
    """K equally long IQ streams through ONE LoRaDemodMultiB200 block.  Returns, per channel, the symbol
    packets, the error signals and the samples consumed, plus the kernel launches of the block."""
    L = _lib()
    L.loopback_rx_multi.argtypes = [C.POINTER(LoopCfg), C.c_void_p, C.c_int, C.c_ulonglong, C.c_ulonglong] + [C.c_void_p] * 5 + \
        [C.POINTER(C.c_ulonglong), C.c_char_p]
    buf = np.ascontiguousarray(streams, np.complex64)
    k, n = buf.shape
    cfg = LoopCfg(sf, 0, 1, 0, 0, mtu, sync, 1.0, 0.0, thresh, 0)
    npk = np.zeros(k, np.int32)
    lens = np.zeros((k, 16), np.int32)
    syms = np.zeros((k, 16, mtu), np.int16)
    errs = np.zeros((k, 16), np.int32)
    cons = np.zeros(k, np.uint64)
    launches = C.c_ulonglong()
    err = C.create_string_buffer(256)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    if L.loopback_rx_multi(C.byref(cfg), vp(buf), k, n, feed, vp(npk), vp(lens), vp(syms), vp(errs), vp(cons), C.byref(launches), err):
        raise RuntimeError(err.value.decode())
    out = []
    for c in range(k):
        out.append(dict(symbols=[syms[c, p, : lens[c, p]].copy() for p in range(min(int(npk[c]), 16))],
                        error=[int(errs[c, p]) for p in range(min(int(npk[c]), 16))], samples=int(cons[c])))
    return out, int(launches.value)


def run(sf, cr, packets, use_b200=False, debug_ports=True, padding=512, mtu=512, sync=0x12, ampl=1.0,
        sigma=0.0, thresh=-30.0, seed=1, feed=0):
    iq = transmit(sf, cr, packets, padding=padding, sync=sync, ampl=ampl, sigma=sigma, seed=seed)
    return receive(sf, cr, iq, use_b200=use_b200, debug_ports=debug_ports, mtu=mtu, sync=sync, thresh=thresh, feed=feed)
