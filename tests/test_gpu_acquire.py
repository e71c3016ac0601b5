"""Row f3 as SURVEY.md section 8 words it: preamble search over ALL candidate offsets in one launch
(lorab200_preamble_search_dev) -- behavioural parity with the reference's sequential hunt
(LoRaDemod.cpp:179-237): on streams built from the unmodified reference encoder + modulator, with
noise and random gaps,
  * exactly one preamble is reported per transmitted packet, in stream order, inside that packet's
    preamble, with the phase the reference's `total = N - value` re-timing (:219) converges to;
  * the CPU oracle's records of the same sliding windows, put through a plain-Python restatement of the
    run rule, give the same list (the kernels add nothing of their own);
  * the unmodified reference block, started on the stream cut at the reported `aligned` offset, posts
    the same packet it posts from the full stream."""
import numpy as np
import pytest

import loopback
from acquire_common import _runs_from_records, _stream

pytestmark = pytest.mark.gpu
needs_chain = pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")


@needs_chain
@pytest.mark.parametrize("sf,stride_div,cfo", [(7, 4, 0.0), (8, 4, 0.4), (10, 4, 0.0), (10, 8, -1.3), (12, 2, 0.0)])
def test_preamble_search_finds_every_packet(lib, cuda_ok, sf, stride_div, cfo):
    import torch
    from oracle import pyoracle

    rng = np.random.default_rng(4000 + 10 * sf + stride_div)
    n = 1 << sf
    n_packets = 4 if sf <= 10 else 2
    x, starts, _ = _stream(sf, rng, n_packets, sigma=0.05, cfo_bins=cfo)
    stride = n // stride_div
    # power - power_avg counts the leakage of an off-bin tone as noise: half a bin of carrier offset leaves
    # -1.7 dB of it on a noise-free chirp (noise windows: -11 dB at SF7, -21 dB at SF10), hence -3 dB
    thresh, min_windows = -3.0, 4 * stride_div          # four symbols of preamble at least
    d = lib.Demod(sf)
    rec, nf = d.preamble_search(torch.from_numpy(x).cuda(), stride=stride, thresh=thresh, min_windows=min_windows)
    assert nf == n_packets == len(rec)
    assert np.all(np.diff(rec["offset"].astype(np.int64)) > 0)          # stream order
    for r, s0 in zip(rec, starts):
        # inside this packet's ten up-chirps, and long enough to be them
        assert s0 - n <= int(r["offset"]) <= s0 + 2 * n
        assert 6 * stride_div <= int(r["windows"]) <= 11 * stride_div
        # the chirp boundary: windows of the reference's own hunt settle at (start - 1 - round(cfo)) mod N
        want = (s0 - 1 - int(np.round(cfo))) % n
        assert (int(r["phase"]) - want) % n in (0, 1, n - 1)
        assert int(r["aligned"]) == int(r["offset"]) + n - ((int(r["offset"]) - int(r["phase"])) % n)
        assert r["snr_db"] > (15.0 if cfo == 0.0 else thresh)

    # the same windows through the CPU oracle and the restated rule
    O = pyoracle.Oracle()
    up, _ = O.tables_demod(sf)
    nw = (x.size - n) // stride + 1
    res, _, _, _ = O.demod_windows(sf, x, nw, up, stride=stride, nthreads=4)
    want = _runs_from_records(res, stride, n, thresh, min_windows)
    got = [(int(r["offset"]), int(r["aligned"]), int(r["phase"]), int(r["windows"])) for r in rec]
    assert got == want

    # fewer record slots than preambles: the count is still the total, the first ones are written
    rec1, nf1 = d.preamble_search(torch.from_numpy(x).cuda(), stride=stride, thresh=thresh, min_windows=min_windows, max_found=1)
    assert nf1 == n_packets and len(rec1) == 1 and rec1[0] == rec[0]
    # nothing to find
    rec0, nf0 = d.preamble_search(torch.from_numpy(x[: starts[0] - 8].copy()).cuda(), stride=stride, thresh=thresh,
                                  min_windows=min_windows)
    assert nf0 == 0 and len(rec0) == 0
    d.close()


@needs_chain
@pytest.mark.parametrize("sf", [7, 9])
def test_reference_receiver_started_at_the_reported_offset(lib, cuda_ok, sf):
    """Behavioural parity: cut the stream at each reported `aligned` offset and give the rest to the UNMODIFIED
    reference chain (demod + decoder) -- its first decoded packet is the payload sent there, as from the full stream."""
    import torch

    rng = np.random.default_rng(7700 + sf)
    n = 1 << sf
    x, starts, payloads = _stream(sf, rng, 3, sigma=0.04)
    d = lib.Demod(sf)
    rec, nf = d.preamble_search(torch.from_numpy(x).cuda(), stride=n // 4, thresh=-3.0, min_windows=16)
    assert nf == 3
    full = loopback.receive(sf, 4, x, use_b200=False, debug_ports=True, mtu=400, thresh=-4.0)
    assert len(full["symbols"]) == 3
    for k, r in enumerate(rec):
        cut = int(r["aligned"])
        assert starts[k] - 1 <= cut <= starts[k] + 3 * n      # (the re-timing settles one sample before the chirp)
        part = loopback.receive(sf, 4, np.ascontiguousarray(x[cut:]), use_b200=False, debug_ports=True, mtu=400, thresh=-4.0)
        assert len(part["symbols"]) == 3 - k
        assert part["decoded"][0] == payloads[k] == full["decoded"][k]
    d.close()
