"""The CPU restatement (oracle/lora_oracle.c) against the REAL reference headers
compiled in place from /root/reference (oracle/_ref/libloraref.so): bit-for-bit.

This is what pins the oracle: kissfft.hh transform,
LoRaDetector.hpp feed/detect, ChirpGenerator.hpp genChirp, the LoRaDemod.cpp
dechirp loop with its fine-tune index recurrence, and -- when the shim was built
with LoRaDemod.cpp -- the block's private table recipes.
"""
import numpy as np
import pytest

from util import bits_equal, noise_windows, synth_windows

SFS = [7, 8, 9, 10, 11, 12]


@pytest.mark.parametrize("sf", SFS)
def test_fft_and_detect_bit_identical(oracle, reference, sf):
    x = noise_windows(sf, 24, seed=100 + sf)
    for w in range(x.shape[0]):
        assert bits_equal(oracle.fft(sf, x[w]), reference.fft(sf, x[w]))
        a = oracle.detect(sf, x[w], want_fft=True)
        b = reference.detect(sf, x[w], want_fft=True)
        assert a[:4] == b[:4]
        assert bits_equal(a[4], b[4])


@pytest.mark.parametrize("sf", SFS)
def test_demod_windows_bit_identical(oracle, reference, sf):
    nw = 40
    iq, _ = synth_windows(sf, nw, seed=sf, snr_db=-3.0)
    iq[3] = 0
    up, down = oracle.tables_demod(sf)
    for chirp in (up, down):
        ro, _, fo, do = oracle.demod_windows(sf, iq, nw, chirp, want_fft=True, want_dec=True, nthreads=2)
        rr, _, fr, dr = reference.demod_windows(sf, iq, nw, chirp, want_fft=True, want_dec=True, nthreads=3)
        assert bits_equal(ro, rr) and bits_equal(fo, fr) and bits_equal(do, dr)


@pytest.mark.parametrize("step", [0.0, 1.28, -3.7, 55.001, 3.001, -0.015625, 127.99])
@pytest.mark.parametrize("sf", [7, 10, 12])
def test_fine_tune_recurrence_bit_identical(oracle, reference, sf, step):
    """LoRaDemod.cpp:157-166: in*chirp*fine[idx] and the int -= float index walk."""
    nw = 12
    iq, _ = synth_windows(sf, nw, seed=sf + 1000, snr_db=0.0)
    up, _ = oracle.tables_demod(sf)
    fine = oracle.table_fine(sf)
    for idx0 in (0, 5, (1 << sf) * 128 - 1):
        ro, io, fo, do = oracle.demod_windows(sf, iq, nw, up, fine=fine, fine_idx=idx0, fine_step=step,
                                              want_fft=True, want_dec=True, nthreads=2)
        rr, ir, fr, dr = reference.demod_windows(sf, iq, nw, up, fine=fine, fine_idx=idx0, fine_step=step,
                                                 want_fft=True, want_dec=True)
        assert io == ir
        assert bits_equal(ro, rr) and bits_equal(fo, fr) and bits_equal(do, dr)


@pytest.mark.parametrize("sf", SFS)
def test_gen_chirp_bit_identical(oracle, reference, sf):
    n = 1 << sf
    for ovs in (1, 2):
        for sym in (0, 1, n // 3, n - 1):
            for down in (False, True):
                f0 = np.float32(2 * np.pi * sym) / np.float32(n * ovs)
                a, pa = oracle.gen_chirp(n, ovs, n * ovs, f0, down, 0.3, np.pi / 4)
                b, pb = reference.gen_chirp(n, ovs, n * ovs, f0, down, 0.3, np.pi / 4)
                assert bits_equal(a, b) and pa == pb
    # quarter chirp with carried phase (LoRaMod.cpp:193)
    a, pa = oracle.gen_chirp(n, 1, n // 4, 0.0, True, 1.0, 1.2345)
    b, pb = reference.gen_chirp(n, 1, n // 4, 0.0, True, 1.0, 1.2345)
    assert bits_equal(a, b) and pa == pb


@pytest.mark.parametrize("sf", SFS)
def test_demod_tables_match_real_block(oracle, reference, sf):
    """Tables as the real LoRaDemod constructor builds them (LoRaDemod.cpp:97-114)."""
    if not reference.has_demod_tables:
        pytest.skip("reference shim built without LoRaDemod.cpp (ministream Pothos surface absent)")
    up, down, fine = reference.ref_tables(sf)
    ou, od = oracle.tables_demod(sf)
    assert bits_equal(up, ou) and bits_equal(down, od)
    assert bits_equal(fine, oracle.table_fine(sf))


@pytest.mark.parametrize("sf", [7, 10, 12])
def test_o3_avx2_courtesy_build_gives_the_same_bits(reference, sf):
    """The -O3 -march=x86-64-v3 build of the reference shim (bench.py's cpu_baseline.o3_avx2 figure)
    must not change a single result bit: no fast-math, no FMA contraction."""
    from oracle import pyoracle
    from util import synth_windows
    try:
        o3 = pyoracle.Reference(o3=True)
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libloraref_o3.so not built")
    nw = 64 if sf <= 10 else 12
    iq, _ = synth_windows(sf, nw, seed=sf + 1, snr_db=-3.0)
    up, _ = reference.tables_demod(sf)
    a, _, fa, _ = reference.demod_windows(sf, iq, nw, up, want_fft=True)
    b, _, fb, _ = o3.demod_windows(sf, iq, nw, up, want_fft=True)
    assert a.tobytes() == b.tobytes()
    assert fa.tobytes() == fb.tobytes()
