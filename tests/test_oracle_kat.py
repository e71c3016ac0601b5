"""The reference's only known-answer test on the hot path, restated without Pothos:
TestDetector.cpp:9-35 -- for every symbol, genChirp up-chirp(sym, phase pi/4) times
genChirp down-chirp through LoRaDetector::detect gives index == sym and power > -10 dB.
The reference runs SF10 only; here SF10 exhaustively and a sample of symbols for the
other SFs, for the oracle, and (when built) for the real reference headers.
"""
import numpy as np
import pytest


def _kat(impl, sf, symbols):
    n = 1 << sf
    down, _ = impl.gen_chirp(n, 1, n, 0.0, True, 1.0, 0.0)
    for sym in symbols:
        chirp, _ = impl.gen_chirp(n, 1, n, np.float32(2 * np.pi * sym) / np.float32(n), False, 1.0,
                                  np.float32(np.pi / 4))
        idx, power, power_avg, f_index = impl.detect(sf, _cmul(down, chirp))
        assert idx == sym, (sf, sym, idx)
        assert power > -10.0


def _cmul(a, b):
    """complex64 multiply with separately rounded products/sums (std::complex<float> operator*)."""
    ar, ai, br, bi = a.real, a.imag, b.real, b.imag
    re = (ar * br).astype(np.float32) - (ai * bi).astype(np.float32)
    im = (ar * bi).astype(np.float32) + (ai * br).astype(np.float32)
    return (re.astype(np.float32) + 1j * im.astype(np.float32)).astype(np.complex64)


def test_detector_kat_sf10_all_symbols(oracle):
    _kat(oracle, 10, range(1024))


@pytest.mark.parametrize("sf", [7, 8, 9, 11, 12])
def test_detector_kat_other_sf(oracle, sf):
    n = 1 << sf
    _kat(oracle, sf, list(range(0, n, max(1, n // 97))) + [n - 1])


def test_detector_kat_real_reference(reference):
    _kat(reference, 10, range(0, 1024, 7))


def test_plan_matches_kissfft_factorisation(oracle):
    """kissfft.hh:34-51: 4s, then 2s; remainders m_s."""
    assert oracle.plan_info(7)[:2] == ([4, 4, 4, 2], [32, 8, 2, 1])
    assert oracle.plan_info(10)[:2] == ([4] * 5, [256, 64, 16, 4, 1])
    assert oracle.plan_info(12)[:2] == ([4] * 6, [1024, 256, 64, 16, 4, 1])
    # LoRaDetector.hpp:18
    assert oracle.plan_info(10)[2] == np.float32(20 * np.log10(1024.0))


def test_detect_edge_cases(oracle):
    """All-zero window: index 0, -inf powers, f_index 0 (demon == 0 branch). Ties: first max wins."""
    idx, p, pa, fi = oracle.detect(7, np.zeros(128, np.complex64))
    assert idx == 0 and np.isneginf(p) and np.isneginf(pa) and fi == 0.0
    # a constant window puts everything in bin 0
    idx, p, pa, fi = oracle.detect(7, np.ones(128, np.complex64))
    assert idx == 0
    # two equal tones -> two equal peaks -> the lower bin wins (strict >)
    n = 128
    i = np.arange(n)
    x = (np.exp(2j * np.pi * 32 * i / n) + np.exp(2j * np.pi * 96 * i / n)).astype(np.complex64)
    idx, *_ = oracle.detect(7, x)
    assert idx in (32, 96)
    r = oracle.detect(7, x, want_fft=True)
    m2 = (r[4].real.astype(np.float32) ** 2 + r[4].imag.astype(np.float32) ** 2)
    assert idx == int(np.argmax(m2))  # np.argmax also returns the first maximum
