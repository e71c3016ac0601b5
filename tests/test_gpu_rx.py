"""Row f3: lorab200_rx_dev -- the reference demod block's whole work() sequence (frame sync,
sync word, down-chirps, quarter chirp, data symbols; LoRaDemod.cpp:145-327) per channel on the
device -- against the UNMODIFIED reference block on the same IQ streams: the same packets with
the same symbols and the same "error" signal, the "power" signal within 1e-4 dB, and the same
number of samples consumed.  Streams come from the reference encoder + modulator with noise and a
random lead-in; a second pass feeds each stream in two pieces and carries the state over."""
import numpy as np
import pytest

import loopback

pytestmark = pytest.mark.gpu
needs_chain = pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")


def _streams(sf, nch, rng, sigma):
    out = []
    for c in range(nch):
        pk = [bytes(rng.integers(0, 256, int(rng.integers(4, 30)), dtype=np.uint8)) for _ in range(3)]
        iq = loopback.transmit(sf, 4, pk, padding=70, sigma=sigma, seed=100 * sf + c)
        lead = int(rng.integers(0, 3 << sf))
        noise = (rng.standard_normal(lead) + 1j * rng.standard_normal(lead)).astype(np.complex64) * np.float32(sigma)
        out.append(np.concatenate([noise, iq]))
    n = max(s.size for s in out)
    buf = np.zeros((nch, n), np.complex64)
    for c, s in enumerate(out):
        buf[c, : s.size] = s
    return buf


def _unpack(r, c):
    npk = int(r["n_packets"][c])
    info = r["info"][c].cpu().numpy()
    syms = r["symbols"][c].cpu().numpy()
    return [dict(symbols=syms[p, : info[p, 0]].copy(), error=int(info[p, 1]),
                 power=float(info[p, 2:3].view(np.float32)[0])) for p in range(npk)]


@needs_chain
@pytest.mark.parametrize("flags", [0, 1], ids=["batch", "simple"])
@pytest.mark.parametrize("sf,thresh", [(7, -30.0), (7, -4.0), (8, -30.0), (9, -30.0), (10, -4.0), (11, -30.0), (12, -30.0), (12, -4.0)])
def test_rx_dev_matches_the_reference_block(lib, cuda_ok, sf, thresh, flags):
    """flags 0: the speculative-batch receiver (rx_batch.cuh); 1 = LORAB200_FLAG_SIMPLE_ONLY: the
    one-window-at-a-time receiver (rx_kernel.cuh) kept as the cross-check."""
    import torch
    rng = np.random.default_rng(sf * 7 + int(-thresh))
    nch, mtu = 5, 96
    buf = _streams(sf, nch, rng, sigma=0.05)
    d = lib.Demod(sf, flags=flags)
    r = d.rx_dev(torch.from_numpy(buf).cuda(), thresh=thresh, mtu=mtu, max_packets=12)
    torch.cuda.synchronize()
    assert d.last_kernel.startswith("rx_simple_sf" if flags else "rx_sf")
    # slots that hold no packet read as empty
    info_all = r["info"].cpu().numpy()
    for c in range(nch):
        assert np.all(info_all[c, int(r["n_packets"][c]):, 0] == 0)
    total_packets = 0
    for c in range(nch):
        ref = loopback.receive(sf, 4, buf[c], use_b200=False, debug_ports=True, mtu=mtu, thresh=thresh)
        got = _unpack(r, c)
        assert len(got) == len(ref["symbols"]), f"channel {c}: {len(got)} packets vs {len(ref['symbols'])}"
        for p, g in enumerate(got):
            assert np.array_equal(g["symbols"], ref["symbols"][p]), f"channel {c} packet {p}"
            assert g["error"] == ref["error"][p]
            assert abs(g["power"] - ref["power"][p]) < 1e-4
        assert int(r["consumed"][c]) == ref["samples"], f"channel {c}: consumed"
        total_packets += len(got)
    assert total_packets >= nch            # every stream carries packets

    # the same streams in two pieces: the state (and the tail the first call left) carries over
    cut = (buf.shape[1] // 2) & ~1
    first = d.rx_dev(torch.from_numpy(np.ascontiguousarray(buf[:, :cut])).cuda(), thresh=thresh, mtu=mtu, max_packets=12)
    torch.cuda.synchronize()
    for c in range(nch):
        used = int(first["consumed"][c])
        rest = np.ascontiguousarray(buf[c: c + 1, used:])
        st = {"state": first["state"]["state"][c: c + 1].clone(), "partial": first["state"]["partial"][c: c + 1].clone()}
        second = d.rx_dev(torch.from_numpy(rest).cuda(), thresh=thresh, mtu=mtu, max_packets=12, state=st)
        torch.cuda.synchronize()
        got = _unpack(first, c) + _unpack(second, 0)
        whole = _unpack(r, c)
        assert len(got) == len(whole)
        for a, b in zip(got, whole):
            assert np.array_equal(a["symbols"], b["symbols"]) and a["error"] == b["error"]
        assert used + int(second["consumed"][0]) == int(r["consumed"][c])
    d.close()


@needs_chain
@pytest.mark.parametrize("sf", [7, 8, 10])
def test_rx_dev_with_carrier_offsets(lib, cuda_ok, sf):
    """Streams with a carrier frequency offset (a fraction of a bin up to a few bins, both signs): the
    preamble then leaves _finefreqError far from zero, positive and negative, and the fine-tune index walks
    the table by tens to hundreds of entries per sample with a wrap every few chunks -- the regime of the
    32-bit region walk, the verified closed-form chunk starts of the one-warp rounds (rx_batch.cuh) and the
    in-region fast step.  Same packets, symbols, error signal and samples consumed as the reference block."""
    import torch
    rng = np.random.default_rng(900 + sf)
    n = 1 << sf
    cfos = [0.31, -0.27, 1.7, -2.4, 0.07, -0.9, 3.3, -0.05]
    streams = []
    for c, cfo in enumerate(cfos):
        pk = [bytes(rng.integers(0, 256, int(rng.integers(6, 24)), dtype=np.uint8)) for _ in range(2)]
        iq = loopback.transmit(sf, 4, pk, padding=40, sigma=0.03, seed=700 * sf + c)
        lead = int(rng.integers(0, 2 << sf))
        noise = (rng.standard_normal(lead) + 1j * rng.standard_normal(lead)).astype(np.complex64) * np.float32(0.03)
        x = np.concatenate([noise, iq])
        rot = np.exp(2j * np.pi * cfo / n * np.arange(x.size)).astype(np.complex64)
        streams.append((x * rot).astype(np.complex64))
    m = max(s.size for s in streams)
    buf = np.zeros((len(streams), m), np.complex64)
    for c, s_ in enumerate(streams):
        buf[c, : s_.size] = s_
    d = lib.Demod(sf)
    for thresh in (-30.0, -6.0):
        r = d.rx_dev(torch.from_numpy(buf).cuda(), thresh=thresh, mtu=64, max_packets=16)
        torch.cuda.synchronize()
        assert d.last_kernel.startswith("rx_sf")
        for c in range(len(streams)):
            ref = loopback.receive(sf, 4, buf[c], use_b200=False, debug_ports=True, mtu=64, thresh=thresh)
            got = _unpack(r, c)
            assert len(got) == len(ref["symbols"]), f"cfo {cfos[c]} thresh {thresh}: {len(got)} packets vs {len(ref['symbols'])}"
            for p_, g in enumerate(got):
                assert np.array_equal(g["symbols"], ref["symbols"][p_]), f"cfo {cfos[c]} packet {p_}"
                assert g["error"] == ref["error"][p_]
            assert int(r["consumed"][c]) == ref["samples"], f"cfo {cfos[c]}: consumed"
    d.close()


@needs_chain
@pytest.mark.parametrize("sf", [7, 9])
def test_iq_to_payload_bytes_entirely_on_the_device(lib, cuda_ok, sf):
    """lorab200_rx_dev -> lorab200_decode_dev: IQ streams in HBM to payload bytes in two launches; the
    bytes equal what the unmodified reference demod + decoder blocks produce from the same streams."""
    import torch
    rng = np.random.default_rng(50 + sf)
    nch, mtu, maxp = 4, 96, 8
    buf = _streams(sf, nch, rng, sigma=0.05)
    d = lib.Demod(sf)
    r = d.rx_dev(torch.from_numpy(buf).cuda(), mtu=mtu, max_packets=maxp)
    lens = r["info"][:, :, 0].contiguous().reshape(-1)                 # symbols per packet slot (0 = unused slot)
    out, blen = d.decode_dev(r["symbols"].reshape(nch * maxp, mtu), lens, rdd=4)
    torch.cuda.synchronize()
    out, blen, npk = out.cpu().numpy(), blen.cpu().numpy(), r["n_packets"].cpu().numpy()
    n_bytes = 0
    for c in range(nch):
        ref = loopback.receive(sf, 4, buf[c], use_b200=False, debug_ports=True, mtu=mtu)
        got = [bytes(out[c * maxp + p, : blen[c * maxp + p]]) for p in range(npk[c]) if blen[c * maxp + p] >= 0]
        assert got == ref["decoded"], f"channel {c}"
        n_bytes += sum(len(g) for g in got)
    assert n_bytes > 0
    d.close()


@needs_chain
@pytest.mark.parametrize("sf,thresh", [(7, -30.0), (9, -4.0), (12, -30.0)])
def test_rx_host_is_bit_identical_to_the_reference_block(lib, cuda_ok, sf, thresh):
    """lorab200_rx_host (what the block's work() calls): host buffers, several channels per call, fed
    in pieces with the unconsumed tail re-presented; packets, symbols, the error signal, the samples
    consumed AND the power signal (recomputed with the host libm from the raw peak / total) equal the
    unmodified reference block's, bit for bit."""
    rng = np.random.default_rng(900 + sf)
    nch, mtu = 3, 96
    buf = _streams(sf, nch, rng, sigma=0.05)
    d = lib.Demod(sf)
    n = 1 << sf
    piece = 11 * n + 7
    for c in range(nch):
        ref = loopback.receive(sf, 4, buf[c], use_b200=False, debug_ports=True, mtu=mtu, thresh=thresh)
        got, state, pos, pending = [], None, 0, np.zeros(0, np.complex64)
        stream = buf[c]
        used_total = 0
        while pos < stream.size or pending.size >= 2 * n:
            pending = np.concatenate([pending, stream[pos:pos + piece]])
            pos += piece
            if pending.size < 2 * n:
                continue
            r = d.rx_host([pending], thresh=thresh, mtu=mtu, max_packets=4, state=state)
            state = r["state"]
            for p in range(int(r["n_packets"][0])):
                info = r["info"][0, p]
                got.append(dict(symbols=r["symbols"][0, p, : info[0]].copy(), error=int(info[1]),
                                power=float(info[2:3].view(np.float32)[0])))
            used = int(r["consumed"][0])
            used_total += used
            pending = pending[used:]
            if used == 0 and pos >= stream.size:
                break
        assert len(got) == len(ref["symbols"]), f"channel {c}: {len(got)} packets vs {len(ref['symbols'])}"
        for p, g in enumerate(got):
            assert np.array_equal(g["symbols"], ref["symbols"][p])
            assert g["error"] == ref["error"][p]
            assert g["power"] == ref["power"][p], "power signal not bit-identical"
        assert used_total == ref["samples"]
    d.close()
