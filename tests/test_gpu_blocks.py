"""The sibling block shells (lora-sdr_b200/host/blocks/LoRaCodecB200.cpp): `/lora/lora_encoder`,
`/lora/lora_mod` and `/lora/lora_decoder` on the GPU, each swapped into the reference chain ONE AT A
TIME (ministream scheduler, oracle/_ref/libloraloop.so) and compared with the unmodified reference
block it replaces -- same setters, same packets in, same packets / samples / drop counts out -- and
finally all four of our blocks (encoder, modulator, demodulator, decoder) as one chain."""
import numpy as np
import pytest

import loopback

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")]

ENC, MOD, DEC = 1, 2, 4


@pytest.fixture(autouse=True)
def _reference_blocks_by_default():
    loopback.use_b200_codec(0)
    yield
    loopback.use_b200_codec(0)


@pytest.mark.parametrize("sf,ppm,rdd,explicit,crc,whitening", [
    (7, 0, 4, 1, 1, 1), (8, 0, 1, 1, 0, 1), (9, 7, 2, 0, 1, 1), (10, 0, 3, 1, 1, 0), (12, 10, 0, 1, 1, 1)])
def test_encoder_block_is_a_drop_in(cuda_ok, sf, ppm, rdd, explicit, crc, whitening):
    """LoRaEncoderB200 vs LoRaEncoder.cpp through the same harness call: symbol packets bit for bit for
    payloads that fill whole interleaver blocks (other lengths: every padding-free block, see
    tests/test_codec.py for the reference's out-of-bounds padding)."""
    rng = np.random.default_rng(sf * 10 + rdd)
    P = ppm or sf
    cfg = (sf, ppm, rdd, explicit, crc, whitening)
    exact = fits = 0
    for ln in list(range(1, 40)) + [55, 89, 144, 233]:
        payload = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
        loopback.use_b200_codec(0)
        ref = loopback.encode_cfg(cfg, payload)
        loopback.use_b200_codec(ENC)
        got = loopback.encode_cfg(cfg, payload)
        assert got.size == ref.size, ln
        n_nib = (ln + 2 * crc) * 2 + 5 * explicit
        fits += n_nib % P == 0
        if n_nib % P == 0:
            assert np.array_equal(got, ref), ln
            exact += 1
        else:
            full = n_nib // P
            keep = 8 + max(0, full - 1) * (4 + rdd) if full >= 1 else 0
            assert np.array_equal(got[:keep], ref[:keep]), ln
    assert exact == fits               # (an odd nibble count never fills blocks of an even PPM)


@pytest.mark.parametrize("sf,rdd,explicit,hdr,crcc,echk,ppm", [
    (7, 4, 1, 0, 0, 0, 0), (7, 1, 1, 1, 1, 1, 0), (9, 2, 0, 0, 1, 0, 7), (10, 3, 1, 1, 0, 0, 0), (12, 4, 0, 0, 0, 1, 10)])
def test_decoder_block_is_a_drop_in(cuda_ok, sf, rdd, explicit, hdr, crcc, echk, ppm):
    """LoRaDecoderB200 vs LoRaDecoder.cpp: clean packets from the reference encoder, corrupted ones,
    truncated ones and garbage -- same length / drop decision / bytes (getDropped included: the harness
    reports a drop as -1)."""
    rng = np.random.default_rng(sf * 100 + rdd * 10 + explicit)
    cfg = (sf, ppm, rdd, explicit, hdr, 9, crcc, echk)
    packets = []
    for _ in range(5):
        s = loopback.encode(sf, rdd, bytes(rng.integers(0, 256, int(rng.integers(2, 40)), dtype=np.uint8)))
        packets.append(s)
        bad = s.copy()
        k = rng.integers(0, s.size, 2)
        bad[k] = (bad[k] + int(rng.integers(1, 1 << sf))) % (1 << sf)
        packets.append(bad)
        packets.append(s[: max(8, s.size - int(rng.integers(1, 6)))])
    for _ in range(6):
        packets.append(rng.integers(0, 1 << sf, int(rng.integers(3, 90))).astype(np.uint16))
    seen = set()
    for p in packets:
        loopback.use_b200_codec(0)
        try:
            exp = loopback.decode(cfg, p)
        except RuntimeError as e:          # the reference's own size_t underflow (tests/test_codec.py)
            assert "bad_alloc" in str(e)
            continue
        loopback.use_b200_codec(DEC)
        got = loopback.decode(cfg, p)
        assert got == exp, (p.size, got[0], exp[0])
        seen.add(exp[0] if exp[0] < 0 else "ok")
    assert len(seen) >= 1


@pytest.mark.parametrize("sf,ovs", [(7, 1), (9, 1), (10, 2), (12, 1)])
def test_mod_block_is_a_drop_in(cuda_ok, sf, ovs):
    """LoRaModB200 vs LoRaMod.cpp: the same number of samples, every sample within 2 float ulp (the
    reference's own phase recurrences evaluated literally; polar() through double sincos), the zero
    padding exact."""
    rng = np.random.default_rng(sf)
    syms = rng.integers(0, 1 << sf, 23).astype(np.uint16)
    loopback.use_b200_codec(0)
    ref = loopback.modulate(sf, syms, padding=3, ampl=0.7, ovs=ovs)
    loopback.use_b200_codec(MOD)
    got = loopback.modulate(sf, syms, padding=3, ampl=0.7, ovs=ovs)
    assert got.size == ref.size
    a, b = got.view(np.float32), ref.view(np.float32)
    ulp = np.spacing(np.maximum(np.abs(b), np.float32(1e-3)))
    assert np.all(np.abs(a - b) <= 2 * ulp)
    assert np.mean(a == b) > 0.9
    tail = 3 * (1 << sf) * ovs
    assert np.all(got[-tail:] == 0) and np.all(ref[-tail:] == 0)


@pytest.mark.parametrize("sf,cr,sigma", [(7, 4, 0.3), (10, 3, 2.0)])
def test_all_four_blocks_are_ours(cuda_ok, sf, cr, sigma):
    """bytes -> LoRaEncoderB200 -> LoRaModB200 -> noise -> LoRaDemodB200 (device receiver) -> LoRaDecoderB200
    -> bytes, on the ministream scheduler; the reference chain on the same payloads decodes the same."""
    rng = np.random.default_rng(40 + sf)
    P = sf
    lens = [ln for ln in range(6, 60) if ((ln + 2) * 2 + 5) % P == 0][:4] or [13, 21]
    pk = [bytes(rng.integers(0, 256, ln, dtype=np.uint8)) for ln in lens]
    loopback.use_b200_codec(0)
    ref = loopback.run(sf, cr, pk, use_b200=False, sigma=sigma, seed=3)
    loopback.use_b200_codec(ENC | MOD | DEC)
    got = loopback.run(sf, cr, pk, use_b200=True, debug_ports=False, sigma=sigma, seed=3)
    assert ref["decoded"] == pk
    assert got["decoded"] == pk
    assert got["dropped"] == ref["dropped"] == 0
    assert got["launches"] > 0 and got["pinned_inputs"] > 0


@pytest.mark.parametrize("sf,thresh,feed_windows", [(7, -30.0, 0), (8, -4.0, 29), (10, -30.0, 0)])
def test_many_channels_behind_one_block(cuda_ok, sf, thresh, feed_windows):
    """LoRaDemodMultiB200: K input ports, K output ports, one launch per work() call; every channel's
    packets, error signals and consumed samples equal those of a separate reference LoRaDemod block fed
    the same stream."""
    rng = np.random.default_rng(70 + sf)
    k, mtu = 6, 96
    streams = []
    for c in range(k):
        pk = [bytes(rng.integers(0, 256, int(rng.integers(4, 30)), dtype=np.uint8)) for _ in range(2)]
        iq = loopback.transmit(sf, 4, pk, padding=70, sigma=0.05, seed=500 * sf + c)
        lead = int(rng.integers(0, 3 << sf))
        noise = (rng.standard_normal(lead) + 1j * rng.standard_normal(lead)).astype(np.complex64) * np.float32(0.05)
        streams.append(np.concatenate([noise, iq]))
    n = max(s.size for s in streams)
    buf = np.zeros((k, n), np.complex64)
    for c, s in enumerate(streams):
        buf[c, : s.size] = s
    feed = feed_windows * (1 << sf) + (3 if feed_windows else 0)
    got, launches = loopback.receive_multi(sf, buf, mtu=mtu, thresh=thresh, feed=feed)
    total = 0
    for c in range(k):
        ref = loopback.receive(sf, 4, buf[c], use_b200=False, debug_ports=True, mtu=mtu, thresh=thresh, feed=feed)
        assert len(got[c]["symbols"]) == len(ref["symbols"]), f"channel {c}"
        for a, b in zip(got[c]["symbols"], ref["symbols"]):
            assert np.array_equal(a, b)
        assert got[c]["error"] == ref["error"]
        assert got[c]["samples"] == ref["samples"]
        total += len(ref["symbols"])
    assert total >= k
    if feed_windows == 0:
        assert launches <= 4              # the whole buffer of every channel in a handful of launches
