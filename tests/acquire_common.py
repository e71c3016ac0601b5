"""Helpers shared by the preamble-search tests (GPU: tests/test_gpu_acquire.py; CPU: tests/test_acquire_rule.py)."""
import numpy as np

import loopback


def _stream(sf, rng, n_packets, sigma, cfo_bins=0.0):
    n = 1 << sf
    parts, starts, pos = [], [], 0
    payloads = []
    for k in range(n_packets):
        gap = int(rng.integers(n // 2, 4 * n)) | 1          # odd gaps: packets start at odd and even offsets
        parts.append(((rng.standard_normal(gap) + 1j * rng.standard_normal(gap)) * sigma).astype(np.complex64))
        pos += gap
        pl = bytes(rng.integers(0, 256, int(rng.integers(6, 20)), dtype=np.uint8))
        payloads.append(pl)
        iq = loopback.transmit(sf, 4, [pl], padding=int(rng.integers(40, 300)), sigma=sigma, seed=31 * sf + k)
        starts.append(pos)
        parts.append(iq)
        pos += iq.size
    x = np.concatenate(parts)
    if cfo_bins:
        x = (x * np.exp(2j * np.pi * cfo_bins / n * np.arange(x.size))).astype(np.complex64)
    return x, starts, payloads


def _runs_from_records(res, stride, n, thresh, min_windows):
    """The run rule of csrc/acquire.cuh restated over an array of detector records."""
    snr = res["power"] - res["power_avg"]
    cand = snr >= thresh
    off = np.arange(res.size, dtype=np.int64) * stride
    phase = (off - res["index"]) % n
    out, w = [], 0
    while w < res.size:
        if not cand[w]:
            w += 1
            continue
        e = w + 1
        while e < res.size and cand[e] and ((phase[e] - phase[e - 1]) % n in (0, 1, n - 1)):
            e += 1
        if e - w >= min_windows:
            out.append((int(off[w]), int(off[w] + n - res["index"][w]), int(phase[w]), e - w))
        w = e
    return out
