"""Shared synthetic-signal helpers for the tests (numpy, seeded)."""
import numpy as np


def ideal_upchirp(sf: int, sym: int, n_windows: int = 1) -> np.ndarray:
    """Closed-form up-chirp whose dechirp with the demod's conj-up table peaks at `sym`.

    Phase is the float64 running sum the demod table recipe integrates
    (LoRaDemod.cpp:97-107) plus a 2*pi*sym*i/N tone.
    """
    n = 1 << sf
    i = np.arange(n, dtype=np.float64)
    # table phase: sum_{k<=i} (-pi + 2*pi*k/N)
    ph = -np.pi * (i + 1) + np.pi * i * (i + 1) / n
    x = np.exp(1j * (ph + 2 * np.pi * sym * i / n))
    return x.astype(np.complex64)


def synth_windows(sf: int, n_windows: int, seed: int, snr_db: float | None = 0.0,
                  amplitude: float = 1.0):
    """n_windows dense windows: random symbol per window + complex AWGN.

    snr_db is the per-sample SNR (signal power amplitude^2, noise power total
    sigma^2); None = no noise.  Returns (iq[n_windows, N] complex64, symbols).
    """
    n = 1 << sf
    rng = np.random.default_rng(seed)
    syms = rng.integers(0, n, size=n_windows)
    i = np.arange(n, dtype=np.float64)
    ph = -np.pi * (i + 1) + np.pi * i * (i + 1) / n
    phase = ph[None, :] + 2 * np.pi * syms[:, None] * i[None, :] / n
    x = amplitude * np.exp(1j * phase)
    if snr_db is not None:
        sigma2 = amplitude ** 2 / (10 ** (snr_db / 10))
        noise = rng.standard_normal((n_windows, n)) + 1j * rng.standard_normal((n_windows, n))
        x = x + noise * np.sqrt(sigma2 / 2)
    return x.astype(np.complex64), syms.astype(np.int64)


def noise_windows(sf: int, n_windows: int, seed: int) -> np.ndarray:
    n = 1 << sf
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n_windows, n)) + 1j * rng.standard_normal((n_windows, n))).astype(np.complex64)


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def same_values(a: np.ndarray, b: np.ndarray) -> bool:
    """Equal as numbers (so +0 == -0), NaNs in the same places."""
    a = np.asarray(a); b = np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
