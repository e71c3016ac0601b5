#!/usr/bin/env python3
"""Generate the golden vectors in this directory FROM THE REAL REFERENCE.

Runs only where /root/reference exists (the build container): it drives the
reference's own headers (LoRaDetector.hpp + kissfft.hh + ChirpGenerator.hpp, and the
LoRaDemod.cpp constructor for the tables) through oracle/_ref/libloraref.so and
records inputs and outputs.  The reference ships no golden vectors of its own
(SURVEY.md section 4), so these recorded outputs are what pins parity on machines
without the reference tree (the GPU box).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

Content per SF (7..12), file sf<SF>.npz:
    iq            [W, N] complex64   seeded input windows (chirp + AWGN, one all-noise, one clean)
    res_plain     [W]    records     feed(i, iq*up[i]) + detect()        (TestDetector.cpp:25 form)
    res_fine0     [W]    records     LoRaDemod.cpp:157-166 form, _finefreqError = 0, idx0 = 0
    res_fine      [W]    records     same with _finefreqError*128 = 1.28, idx0 = 5; fine_end = index after
    fft_plain     [W, N] complex64   FFT bins of res_plain (SF <= 9) or their sha256 (SF >= 10)
    up/down/twiddle/fine sha256      table fingerprints (LoRaDemod.cpp:97-114, kissfft.hh:17-22)
    genchirp      [N]    complex64   genChirp(N,1,N, f0(sym=N/3), up, ampl 0.3, phase pi/4), phase_out
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import pyoracle  # noqa: E402
from util import noise_windows, synth_windows  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = pyoracle.Reference()
    assert ref.has_demod_tables, "build oracle/_ref with the ministream Pothos surface to record the real tables"
    out = Path(__file__).resolve().parent
    for sf in range(7, 13):
        n = 1 << sf
        W = 6
        iq, syms = synth_windows(sf, W, seed=0x60D + sf, snr_db=-6.0)
        iq[1] = noise_windows(sf, 1, seed=sf)[0]
        iq[2], _ = ref.gen_chirp(n, 1, n, np.float32(2 * np.pi * 5) / np.float32(n), False, 1.0, 0.0)  # clean genChirp
        up, down, fine = ref.ref_tables(sf)
        tw = ref.twiddles(sf)
        r_plain, _, fft, _ = ref.demod_windows(sf, iq, W, up, want_fft=True)
        r_f0, e0, _, _ = ref.demod_windows(sf, iq, W, up, fine=fine, fine_idx=0, fine_step=0.0)
        r_f, e1, _, _ = ref.demod_windows(sf, iq, W, up, fine=fine, fine_idx=5, fine_step=1.28)
        gc, gph = ref.gen_chirp(n, 1, n, np.float32(2 * np.pi * (n // 3)) / np.float32(n), False, 0.3,
                                np.float32(np.pi / 4))
        d = dict(iq=iq, syms=syms, res_plain=r_plain, res_fine0=r_f0, res_fine=r_f,
                 fine_end=np.array([e0, e1]), genchirp=gc, genchirp_phase=np.float32(gph),
                 sha_up=sha(up), sha_down=sha(down), sha_fine=sha(fine), sha_twiddle=sha(tw),
                 sha_fft=sha(fft))
        if sf <= 9:
            d["fft_plain"] = fft
        np.savez_compressed(out / f"sf{sf}.npz", **d)
        print(f"sf{sf}: indices {r_plain['index'].tolist()}  sent {syms.tolist()}")


if __name__ == "__main__":
    main()
