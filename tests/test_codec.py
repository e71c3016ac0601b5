"""Row f4 on the CPU: the product's symbols -> bytes codec (lora-sdr_b200/csrc/codec.h, built for
the host by oracle/codec_host.cpp) against the UNMODIFIED reference LoRaDecoder.cpp run through
the ministream harness -- every setter combination, clean packets from the reference encoder,
packets with symbol errors, truncated packets and plain garbage.  Lengths, drop decisions and
bytes must be identical.  (The GPU test, tests/test_gpu_decode.py, then checks the device build
of the same header against both.)"""
import ctypes as C
import itertools
import subprocess
from pathlib import Path

import numpy as np
import pytest

import loopback

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "libcodec_host.so"

needs_chain = pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")


def host_codec():
    if not LIB.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "libcodec_host.so"], check=True, capture_output=True)
    L = C.CDLL(str(LIB))
    L.codec_host_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return L


def host_decode(L, cfg8, syms):
    cfg = (C.c_int32 * 8)(*[int(v) for v in cfg8])
    syms = np.ascontiguousarray(syms, np.uint16)
    out = np.zeros(8192, np.uint8)
    n = L.codec_host_decode(cfg, syms.ctypes.data_as(C.c_void_p), syms.size, out.ctypes.data_as(C.c_void_p), out.size)
    return n, bytes(out[: max(n, 0)])


def configs():
    """(sf, ppm, rdd, explicit, hdr, data_length, crcc, error_check)"""
    for sf, rdd in itertools.product((7, 9, 12), range(5)):
        for ppm in (0, sf - 2):
            for explicit, hdr, crcc, echk in itertools.product((1, 0), (0, 1), (0, 1), (0, 1)):
                if not explicit and hdr:
                    continue
                yield (sf, ppm, rdd, explicit, hdr, 11, crcc, echk)


@needs_chain
def test_codec_matches_reference_decoder_on_encoder_output_and_corruptions():
    L = host_codec()
    rng = np.random.default_rng(2024)
    checked = dropped = 0
    for cfg in configs():
        sf, ppm, rdd, explicit, hdr, dlen, crcc, echk = cfg
        if ppm:
            continue                       # the driver's encoder entry point has no symbol-size setter
        payload = bytes(rng.integers(0, 256, dlen, dtype=np.uint8))
        syms = loopback.encode(sf, rdd, payload)            # explicit header, no CRC, rate rdd
        variants = [syms]
        for _ in range(3):                                  # symbol errors (off-by-one and random)
            s = syms.copy()
            k = rng.integers(0, s.size, rng.integers(1, 4))
            s[k] = (s[k] + rng.choice([1, (1 << sf) - 1, int(rng.integers(1, 1 << sf))])) % (1 << sf)
            variants.append(s)
        variants.append(syms[: max(8, syms.size - 3)])      # truncated
        variants.append(syms[:5])                           # shorter than a header: ignored
        for s in variants:
            exp = loopback.decode(cfg, s)
            got = host_decode(L, cfg, s)
            assert got == exp, (cfg, list(s[:12]), got, exp)
            checked += 1
            dropped += exp[0] == -1
        if explicit and not crcc and not hdr:
            assert loopback.decode(cfg, syms)[1][: dlen - 2] == payload[: dlen - 2]   # (the reference trims 5 bytes)
    assert checked > 500 and dropped > 10


@needs_chain
def test_codec_matches_reference_decoder_on_random_symbols():
    """Garbage in: header fields, lengths, CRC and parity paths are all exercised by random symbols."""
    L = host_codec()
    rng = np.random.default_rng(7)
    outcomes = {"ok": 0, "dropped": 0, "ignored": 0}
    for cfg in configs():
        sf = cfg[0]
        for n in (8, 13, 40, 83):
            s = rng.integers(0, 1 << sf, n).astype(np.uint16)
            got = host_decode(L, cfg, s)
            try:
                exp = loopback.decode(cfg, s)
            except RuntimeError as e:
                # explicit header, hdr off, decoded length < 2 (or 5 with CRC): the reference computes
                # dataLength -= 5 on a size_t and dies allocating the result (LoRaDecoder.cpp:381-384,393);
                # ours reports the underflow as "no room"
                assert "bad_alloc" in str(e) and got[0] == -3, (cfg, n, got[0], str(e))
                outcomes["underflow"] = outcomes.get("underflow", 0) + 1
                continue
            assert got == exp, (cfg, n, got[0], exp[0])
            outcomes["ok" if exp[0] >= 0 else ("dropped" if exp[0] == -1 else "ignored")] += 1
    assert outcomes["ok"] > 100 and outcomes["dropped"] > 100


def host_encode(L, cfg6, payload):
    L.codec_host_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    cfg = (C.c_int32 * 6)(*[int(v) for v in cfg6])
    out = np.zeros(4096, np.uint16)
    n = L.codec_host_encode(cfg, payload, len(payload), out.ctypes.data_as(C.c_void_p), out.size)
    return out[:n].copy()


def enc_configs():
    """(sf, ppm, rdd, explicit, crc, whitening)"""
    for sf, rdd, ex, crc, wh in itertools.product((7, 9, 12), range(5), (1, 0), (1, 0), (1, 0)):
        for ppm in (0, sf - 2):
            yield (sf, ppm, rdd, ex, crc, wh)


@needs_chain
def test_encoder_matches_reference_encoder():
    """The product's encode_packet against the unmodified LoRaEncoder.cpp for every setter combination.
    Payload lengths whose codewords fill whole interleaver blocks compare bit for bit; other lengths
    (the reference pads with bytes read past the end of its vector, i.e. unspecified ones) must agree on
    the symbol count and on every block that contains no padding, and must decode back to the payload."""
    L = host_codec()
    rng = np.random.default_rng(99)
    exact = padded = 0
    for cfg in enc_configs():
        sf, ppm, rdd, ex, crc, wh = cfg
        P = ppm or sf
        for ln in (1, 2, 5, 9, 16, 23, 31, 40):
            payload = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
            ref = loopback.encode_cfg(cfg, payload)
            got = host_encode(L, cfg, payload)
            assert got.size == ref.size, (cfg, ln)
            n_nib = (ln + 2 * crc) * 2 + 5 * ex
            if n_nib % P == 0:
                assert np.array_equal(got, ref), (cfg, ln)
                exact += 1
            else:
                full_blocks = n_nib // P                      # blocks without padding
                keep = 8 + max(0, full_blocks - 1) * (4 + rdd) if full_blocks >= 1 else 0
                assert np.array_equal(got[:keep], ref[:keep]), (cfg, ln)
                padded += 1
            if ex and wh and ln >= 2:
                n, out = host_decode(L, (sf, ppm, rdd, 1, 0, 8, 0, 0), got)
                assert out[:ln] == payload[:ln] if crc else out[: ln - 2] == payload[: ln - 2], (cfg, ln)
    for ln in range(1, 60):                                    # denser sweep of exact-fit lengths
        for cfg in ((7, 0, 4, 1, 1, 1), (10, 0, 1, 1, 0, 1), (12, 10, 3, 0, 1, 1)):
            P = cfg[1] or cfg[0]
            if ((ln + 2 * cfg[4]) * 2 + 5 * cfg[3]) % P == 0:
                payload = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
                assert np.array_equal(host_encode(L, cfg, payload), loopback.encode_cfg(cfg, payload))
                exact += 1
    assert exact > 100 and padded > 500


def test_codec_round_trip_and_single_error_correction():
    """Size-independent properties of the product codec (host build, no reference needed):
    decode(encode(p)) == p for every configuration, and with coding rates 4/7 and 4/8 any single
    corrupted bit per interleaver block is corrected (one symbol bit lands in exactly one codeword)."""
    L = host_codec()
    rng = np.random.default_rng(12345)
    for trial in range(400):
        sf = int(rng.integers(7, 13))
        ppm = int(rng.choice([0, sf - 1, sf - 2]))
        rdd = int(rng.integers(0, 5))
        crc = int(rng.integers(0, 2))
        wh = int(rng.integers(0, 2))
        ln = int(rng.integers(2, 80))
        payload = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
        syms = host_encode(L, (sf, ppm, rdd, 1, crc, wh), payload)
        if not wh:
            continue                                        # the decoder always de-whitens (LoRaDecoder.cpp:229-258)
        dec_cfg = (sf, ppm, rdd, 1, 0, 8, crc, 1)           # error check on: a bad parity would drop the packet
        n, out = host_decode(L, dec_cfg, syms)
        want = payload if crc else payload[: ln - 2]        # (hdr off trims 5 bytes either way, see codec.h)
        assert n == len(want) and out == want, (sf, ppm, rdd, crc, ln)
        if rdd >= 3:
            bad = syms.copy()
            P = ppm or sf
            shift = sf - P
            # one flipped (significant) bit in one symbol of every block
            blocks = [(0, 8)] + [(8 + b * (4 + rdd), 4 + rdd) for b in range((syms.size - 8) // (4 + rdd))]
            for start, count in blocks:
                k = start + int(rng.integers(0, count))
                g = int(bad[k]) >> shift
                g ^= g >> 1                                  # to Gray, flip one bit there, back to binary
                g ^= 1 << int(rng.integers(0, P))
                b = g
                for s in (1, 2, 4, 8):
                    b ^= b >> s
                bad[k] = (b << shift) & 0xFFFF
            n2, out2 = host_decode(L, (sf, ppm, rdd, 1, 0, 8, crc, 0), bad)
            assert n2 == len(want) and out2 == want, ("corrected", sf, ppm, rdd, crc, ln)


@needs_chain
@pytest.mark.parametrize("sf", [8, 10, 12])
def test_one_block_explicit_packets_stay_inside_the_scratch(sf):
    """Explicit header, rdd = 4, even PPM, exactly 8 symbols: the first block's (8,4) nibbles end on
    byte n_cw / 2 (the reference writes one byte past its vector there, LoRaDecoder.cpp:315-340).
    codec_host_decode sizes both scratch areas exactly with a guard byte behind each and returns
    -100 if a guard was touched."""
    L = host_codec()
    rng = np.random.default_rng(sf)
    for ppm in (0, sf - 2):
        for hdr, crcc, echk in itertools.product((0, 1), (0, 1), (0, 1)):
            cfg = (sf, ppm, 4, 1, hdr, 11, crcc, echk)
            for _ in range(6):
                syms = rng.integers(0, 1 << sf, 8).astype(np.uint16)
                n, data = host_decode(L, cfg, syms)
                assert n != -100, ("scratch overrun", cfg)
                exp_len, exp = loopback.decode(cfg, syms)
                assert n == exp_len and (n <= 0 or data == exp), (cfg, n, exp_len)
