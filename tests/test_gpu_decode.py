"""Row f4 on the GPU: lorab200_decode_dev (one packet per thread, the device build of codec.h)
against the unmodified reference LoRaDecoder.cpp: a batch that mixes clean packets from the
reference encoder, corrupted ones, truncated ones and random symbols, with per-packet lengths;
every length / drop decision / byte must be identical.  Then the whole receive chain on the
device: reference-encoded packets -> GPU modulator -> GPU demod -> GPU decoder -> payload."""
import numpy as np
import pytest

import loopback

pytestmark = pytest.mark.gpu
needs_chain = pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")


def _batch(sf, rdd, rng, n_clean=6, n_rand=10):
    packets = []
    for _ in range(n_clean):
        payload = bytes(rng.integers(0, 256, int(rng.integers(2, 40)), dtype=np.uint8))
        s = loopback.encode(sf, rdd, payload)
        packets.append(s)
        bad = s.copy()
        k = rng.integers(0, s.size, 2)
        bad[k] = (bad[k] + int(rng.integers(1, 1 << sf))) % (1 << sf)
        packets.append(bad)
        packets.append(s[: max(8, s.size - int(rng.integers(1, 6)))])
    for _ in range(n_rand):
        packets.append(rng.integers(0, 1 << sf, int(rng.integers(3, 90))).astype(np.uint16))
    return packets


@needs_chain
@pytest.mark.parametrize("sf,rdd,explicit,hdr,crcc,echk,ppm", [
    (7, 4, 1, 0, 0, 0, 0), (7, 1, 1, 1, 1, 1, 0), (9, 0, 1, 0, 0, 1, 0), (9, 2, 0, 0, 1, 0, 7),
    (10, 3, 1, 1, 0, 0, 0), (12, 4, 0, 0, 0, 1, 10), (12, 2, 1, 0, 1, 0, 0), (8, 3, 0, 0, 0, 0, 0)])
def test_decode_dev_matches_reference_decoder(lib, cuda_ok, sf, rdd, explicit, hdr, crcc, echk, ppm):
    import torch
    rng = np.random.default_rng(sf * 100 + rdd * 10 + explicit)
    packets = _batch(sf, rdd, rng)
    stride = max(p.size for p in packets)
    host = np.zeros((len(packets), stride), np.uint16)
    lens = np.zeros(len(packets), np.int32)
    for i, p in enumerate(packets):
        host[i, : p.size] = p
        host[i, p.size:] = 0xFFFF                       # garbage after the valid symbols must be ignored
        lens[i] = p.size
    cfg = (sf, ppm, rdd, explicit, hdr, 9, crcc, echk)
    d = lib.Demod(sf)
    out, got_len = d.decode_dev(torch.from_numpy(host.view(np.int16)).cuda(), torch.from_numpy(lens).cuda(),
                                ppm=ppm, rdd=rdd, explicit_header=bool(explicit), hdr=bool(hdr), data_length=9,
                                crcc=bool(crcc), error_check=bool(echk), bytes_stride=512)
    assert d.last_kernel == "decode"
    out, got_len = out.cpu().numpy(), got_len.cpu().numpy()
    seen = set()
    for i, p in enumerate(packets):
        try:
            exp_len, exp = loopback.decode(cfg, p)
        except RuntimeError as e:                       # the reference's own size_t underflow (see tests/test_codec.py)
            assert "bad_alloc" in str(e) and got_len[i] == -3
            continue
        assert got_len[i] == exp_len, (i, p.size, got_len[i], exp_len)
        if exp_len > 0:
            assert bytes(out[i, :exp_len]) == exp
        seen.add("ok" if exp_len >= 0 else exp_len)
    if explicit and not crcc:
        assert "ok" in seen            # (with the CRC check on, packets made without a CRC are all dropped)
    assert len(seen) >= 1
    d.close()


@needs_chain
@pytest.mark.parametrize("sf", [7, 10])
def test_whole_receive_chain_on_the_device(lib, cuda_ok, sf):
    """reference encoder -> lorab200_mod_dev (ideal) -> lorab200_demod_dev on the data symbols ->
    lorab200_decode_dev: the payloads come back, nothing leaves the GPU in between."""
    import torch
    rng = np.random.default_rng(sf)
    n = 1 << sf
    payloads = [bytes(rng.integers(0, 256, 20, dtype=np.uint8)) for _ in range(8)]
    syms = np.stack([loopback.encode(sf, 4, p) for p in payloads])
    d = lib.Demod(sf)
    iq = d.mod_dev(torch.from_numpy(syms.astype(np.int16)).cuda(), padding=1, ampl=1.0, mode=lib.MOD_IDEAL)
    nsy = syms.shape[1]
    head = 14 * n + n // 4                                   # preamble, sync, down-chirps, quarter chirp
    data = iq[:, head: head + nsy * n].contiguous()
    res, _, _, _ = d.demod_dev(data.reshape(-1), len(payloads) * nsy)
    torch.cuda.synchronize()
    idx = lib.results_to_numpy(res)["index"].reshape(len(payloads), nsy)
    # without the receiver's sync stage the detected bin is the symbol plus one constant (the modulator's
    # chirp law starts one frequency step in, ChirpGenerator.hpp:33 "f += fStep" before the first sample)
    off = int(idx[0, 0] - syms[0, 0]) % n
    assert off in (0, 1) and np.array_equal((idx - off) % n, syms)
    idx = (idx - off) % n
    out, lens = d.decode_dev(torch.from_numpy(idx.astype(np.int16)).cuda(), rdd=4)
    out, lens = out.cpu().numpy(), lens.cpu().numpy()
    for i, p in enumerate(payloads):
        ref_len, ref = loopback.decode((sf, 0, 4, 1, 0, 8, 0, 0), syms[i])
        assert lens[i] == ref_len and bytes(out[i, : lens[i]]) == ref
        assert ref[: len(p) - 2] == p[: len(p) - 2]
    d.close()


@needs_chain
@pytest.mark.parametrize("sf,ppm,rdd,explicit,crc,whitening", [
    (7, 0, 4, 1, 1, 1), (8, 0, 1, 1, 0, 1), (9, 7, 2, 0, 1, 1), (10, 0, 3, 1, 1, 0), (12, 10, 0, 1, 1, 1), (12, 0, 4, 0, 0, 1)])
def test_encode_dev_matches_reference_encoder(lib, cuda_ok, sf, ppm, rdd, explicit, crc, whitening):
    """lorab200_encode_dev against the unmodified LoRaEncoder.cpp: exact-fit payload lengths bit for bit,
    other lengths on every padding-free block (tests/test_codec.py explains the reference's padding)."""
    import torch
    rng = np.random.default_rng(sf + rdd)
    P = ppm or sf
    lens = np.array([1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233] + list(range(4, 40)), np.int32)
    stride = int(lens.max())
    host = rng.integers(0, 256, (lens.size, stride), dtype=np.uint8)
    d = lib.Demod(sf)
    syms, nsy = d.encode_dev(torch.from_numpy(host).cuda(), torch.from_numpy(lens).cuda(), ppm=ppm, rdd=rdd,
                             explicit_header=bool(explicit), crc=bool(crc), whitening=bool(whitening))
    assert d.last_kernel == "encode"
    syms, nsy = syms.cpu().numpy().view(np.uint16), nsy.cpu().numpy()
    exact = 0
    for i, ln in enumerate(lens):
        ref = loopback.encode_cfg((sf, ppm, rdd, explicit, crc, whitening), bytes(host[i, :ln]))
        assert nsy[i] == ref.size
        n_nib = (int(ln) + 2 * crc) * 2 + 5 * explicit
        if n_nib % P == 0:
            assert np.array_equal(syms[i, : ref.size], ref), (i, ln)
            exact += 1
        else:
            full = n_nib // P
            keep = 8 + max(0, full - 1) * (4 + rdd) if full >= 1 else 0
            assert np.array_equal(syms[i, :keep], ref[:keep]), (i, ln)
    fits = sum(((int(ln) + 2 * crc) * 2 + 5 * explicit) % P == 0 for ln in lens)
    assert exact == fits               # (an odd nibble count never fills blocks of an even PPM)
    d.close()


@needs_chain
def test_transmit_and_receive_chains_on_the_device(lib, cuda_ok):
    """bytes -> encode_dev -> mod_dev -> rx_dev -> decode_dev -> bytes, four launches, no reference code in
    the loop; the reference receiver decodes the same frames to the same bytes."""
    import torch
    sf, nch, ln = 8, 6, 17
    rng = np.random.default_rng(8)
    payloads = rng.integers(0, 256, (nch, ln), dtype=np.uint8)
    d = lib.Demod(sf)
    syms, nsy = d.encode_dev(torch.from_numpy(payloads).cuda(), rdd=4, crc=True)
    assert int(nsy.min()) == int(nsy.max()) == syms.shape[1]
    iq = d.mod_dev(syms, padding=140, ampl=1.0, mode=lib.MOD_IDEAL)
    iq = iq + 0.02 * torch.view_as_complex(torch.randn((*iq.shape, 2), device=iq.device, generator=torch.Generator("cuda").manual_seed(1)))
    r = d.rx_dev(iq.contiguous(), mtu=128, max_packets=4)
    lens = r["info"][:, :, 0].contiguous().reshape(-1)
    out, blen = d.decode_dev(r["symbols"].reshape(nch * 4, 128), lens, rdd=4)
    torch.cuda.synchronize()
    out, blen, npk = out.cpu().numpy(), blen.cpu().numpy(), r["n_packets"].cpu().numpy()
    host_iq = iq.cpu().numpy()
    for c in range(nch):
        assert npk[c] == 1
        assert bytes(out[c * 4, :ln]) == bytes(payloads[c])
        ref = loopback.receive(sf, 4, host_iq[c], use_b200=False, debug_ports=True, mtu=128)
        assert ref["decoded"] and ref["decoded"][0][:ln] == bytes(payloads[c])
    d.close()


@needs_chain
@pytest.mark.parametrize("sf", [8, 10, 12])
def test_decode_dev_one_block_packets_packed_at_stride_8(lib, cuda_ok, sf):
    """A batch of 8-symbol explicit packets at symbols_stride == 8 (even PPM, rdd 4): packet p's first
    block ends on the byte behind its payload scratch; it must not land in packet p+1's header byte
    (per-packet scratch = codec_scratch_bytes).  Every packet must decode as it does alone."""
    import torch
    rng = np.random.default_rng(sf + 77)
    n_pk = 64
    clean = loopback.encode(sf, 4, bytes([1, 2, 3]))[:8]        # a valid header block of a short packet
    host = rng.integers(0, 1 << sf, (n_pk, 8)).astype(np.uint16)
    host[::3] = clean
    lens = np.full(n_pk, 8, np.int32)
    cfg = (sf, 0, 4, 1, 1, 9, 0, 0)
    d = lib.Demod(sf)
    out, got_len = d.decode_dev(torch.from_numpy(host.view(np.int16)).cuda(), torch.from_numpy(lens).cuda(),
                                ppm=0, rdd=4, explicit_header=True, hdr=True, data_length=9, crcc=False,
                                error_check=False, bytes_stride=64)
    out, got_len = out.cpu().numpy(), got_len.cpu().numpy()
    for i in range(n_pk):
        exp_len, exp = loopback.decode(cfg, host[i])
        assert got_len[i] == exp_len, (i, got_len[i], exp_len)
        if exp_len > 0:
            assert bytes(out[i, :exp_len]) == exp, i
    d.close()
