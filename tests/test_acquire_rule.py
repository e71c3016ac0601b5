"""The preamble-search RULE (csrc/acquire.cuh) without a GPU: the CPU oracle's detector records of the sliding
windows of a stream, put through the plain-Python restatement of the run rule, give exactly one preamble per
transmitted packet, at the phase the reference's `total = N - value` re-timing (LoRaDemod.cpp:219) settles on.
(The GPU test checks that the kernels return the same list.)"""
import numpy as np
import pytest

import loopback
from acquire_common import _runs_from_records, _stream

needs_chain = pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")


@needs_chain
@pytest.mark.parametrize("sf,stride_div,cfo", [(7, 4, 0.0), (7, 8, 0.5), (8, 4, -0.4), (9, 2, 2.2)])
def test_run_rule_finds_every_preamble(sf, stride_div, cfo):
    from oracle import pyoracle

    rng = np.random.default_rng(1234 + 10 * sf + stride_div)
    n = 1 << sf
    x, starts, _ = _stream(sf, rng, 3, sigma=0.05, cfo_bins=cfo)
    stride = n // stride_div
    O = pyoracle.Oracle()
    up, _ = O.tables_demod(sf)
    nw = (x.size - n) // stride + 1
    res, _, _, _ = O.demod_windows(sf, x, nw, up, stride=stride, nthreads=4)
    runs = _runs_from_records(res, stride, n, -3.0, 4 * stride_div)
    assert len(runs) == len(starts)
    for (off, aligned, phase, windows), s0 in zip(runs, starts):
        assert s0 - n <= off <= s0 + 2 * n
        assert 6 * stride_div <= windows <= 11 * stride_div
        want = (s0 - 1 - int(np.round(cfo))) % n
        assert (phase - want) % n in (0, 1, n - 1)
        assert aligned % n == phase % n or (aligned - phase) % n in (0, 1, n - 1)
    # noise alone: nothing
    noise = ((rng.standard_normal(20 * n) + 1j * rng.standard_normal(20 * n)) * 0.05).astype(np.complex64)
    nw = (noise.size - n) // stride + 1
    res, _, _, _ = O.demod_windows(sf, noise, nw, up, stride=stride)
    assert _runs_from_records(res, stride, n, -3.0, 4 * stride_div) == []
