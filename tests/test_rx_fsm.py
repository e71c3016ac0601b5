"""The receiver's state machine (lora-sdr_b200/csrc/rx_fsm.h, the restatement of
LoRaDemod.cpp:173-327 that lorab200_rx_dev runs per channel) on the CPU: driven window by window
with the oracle's detector outputs, it must post the same symbol packets, emit the same `error`
signal and consume the same number of samples as the UNMODIFIED reference block on the same
stream.  (tests/test_gpu_rx.py checks the device kernel against the reference block directly.)"""
import ctypes as C

import numpy as np
import pytest

import loopback
from test_codec import host_codec

needs_chain = pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")


class Result(C.Structure):
    _fields_ = [("index", C.c_int32), ("power", C.c_float), ("power_avg", C.c_float), ("f_index", C.c_float)]


def run_fsm(L, oracle, tables, sf, iq, sync, thresh, mtu):
    n = 1 << sf
    up, down, fine = tables
    cfg = (C.c_int32 * 4)(n, sync, mtu, int(np.float32(thresh).view(np.int32)))
    state = (C.c_int32 * 14)()                         # all zero = activate()
    L.rx_host_needs_second.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.rx_host_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    off, packets, errors, cur = 0, [], [], []

    def demod(at, table, idx, ffe):
        res, idx_end, _, _ = oracle.demod_windows(sf, iq[at: at + n], 1, down if table == 1 else up, fine=fine,
                                                  fine_idx=idx, fine_step=float(np.float32(ffe) * np.float32(128.0)))
        r = Result(int(res["index"][0]), float(res["power"][0]), float(res["power_avg"][0]), float(res["f_index"][0]))
        return r, idx_end

    while iq.size - off >= 2 * n:
        table, fine_index = state[1], state[5]
        ffe = float(np.array([state[6]], np.int32).view(np.float32)[0])
        r0, after0 = demod(off, table, fine_index, ffe)
        need2 = L.rx_host_needs_second(cfg, state, C.byref(r0))
        r1 = Result()
        if need2:
            r1, _ = demod(off + n, table, after0, ffe)
        phase_before = state[0]
        out = (C.c_int64 * 3)()
        L.rx_host_step(cfg, state, C.byref(r0), after0, need2, C.byref(r1), out)
        if phase_before == 2:                           # DOWNCHIRP1 just ran: the "error" signal
            errors.append(state[7])
        if out[1] >= 0:
            cur.append(out[1])
        if out[2]:
            packets.append(np.array(cur, np.int16))
            cur = []
        off += out[0]
    return packets, errors, off


@needs_chain
@pytest.mark.parametrize("sf,thresh", [(7, -30.0), (7, -4.0), (8, -30.0), (9, -4.0)])
def test_state_machine_matches_the_reference_block(oracle, sf, thresh):
    L = host_codec()
    rng = np.random.default_rng(sf)
    up, down = oracle.tables_demod(sf)
    tables = (up, down, oracle.table_fine(sf))
    pk = [bytes(rng.integers(0, 256, int(rng.integers(4, 20)), dtype=np.uint8)) for _ in range(3)]
    iq = loopback.transmit(sf, 4, pk, padding=40, sigma=0.05, seed=sf)
    lead = (rng.standard_normal(300) + 1j * rng.standard_normal(300)).astype(np.complex64) * np.float32(0.05)
    iq = np.concatenate([lead, iq])
    ref = loopback.receive(sf, 4, iq, use_b200=False, debug_ports=True, mtu=64, thresh=thresh)
    packets, errors, consumed = run_fsm(L, oracle, tables, sf, iq, 0x12, thresh, 64)
    assert len(packets) == len(ref["symbols"]) and len(packets) >= 3
    for a, b in zip(packets, ref["symbols"]):
        assert np.array_equal(a, b)
    assert errors == ref["error"][: len(errors)] and len(errors) >= len(packets)
    assert consumed == ref["samples"]
