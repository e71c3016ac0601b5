"""SURVEY.md section 8(d) strawman: the UNFUSED GPU pipeline a library user would write --
elementwise dechirp -> batched cuFFT (torch.fft) -> |.|^2 -> argmax + sum -- timed with the
same synthetic batches, buffer rotation and CUDA-event method as bench.py.  It moves the
window through HBM about five times (read IQ, write/read the dechirped copy, write/read the
spectrum, read |.|^2) where the fused kernel reads it once, and its results are not
bit-identical to the reference (cuFFT's factorisation is not kissfft's).  One JSON line.

    python tools/strawman_cufft.py --sf 7 --windows 65536 --steps 200
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import lora_sdr_b200 as m  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=int, default=7)
    ap.add_argument("--windows", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    m.build()
    sf, nw, n = a.sf, a.windows, 1 << a.sf
    d = m.Demod(sf)
    chirp = torch.from_numpy(d.table(m.TABLE_UP)).to(dev)
    step_bytes = nw * n * 8
    nbuf = min(64, max(2, int(np.ceil(3 * bench.L2_BYTES / step_bytes)) + 1))
    bufs, syms = [], []
    for b in range(nbuf):
        iq, s = bench.synth_on_device(sf, nw, seed=0x10AA + 7919 * b, device=dev)
        bufs.append(iq.view(nw, n))
        syms.append(s)

    def step(iq):
        x = iq * chirp[None, :]
        f = torch.fft.fft(x, dim=1)
        p = f.real * f.real + f.imag * f.imag
        return p.argmax(dim=1), p.sum(dim=1, dtype=torch.float64)

    for i in range(max(3, a.warmup)):
        idx, tot = step(bufs[i % nbuf])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        idx, tot = step(bufs[i % nbuf])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    last = (a.steps - 1) % nbuf
    ours, *_ = d.demod_dev(bufs[last].reshape(-1), nw)
    torch.cuda.synchronize()
    ours_idx = torch.from_numpy(m.results_to_numpy(ours)["index"].astype(np.int64)).to(dev)
    value = nw * a.steps / (ms * 1e-3) / 1e6
    print(json.dumps({
        "impl": "cufft_strawman", "metric": "Msymbols/s demod (dechirp+FFT+argmax)", "value": value,
        "unit": "Msymbols/s", "ms_per_step": ms / a.steps, "steps": a.steps,
        "config": {"workload": bench.workload_name(sf, nw), "sf": sf, "windows_per_step": nw,
                   "pipeline": "torch mul -> torch.fft.fft (cuFFT C2C batched) -> re^2+im^2 -> argmax, f64 sum"},
        "achieved_algorithmic_GBps": nw * (8 * n + 16) * a.steps / (ms * 1e-3) / 1e9,
        "index_agreement_with_fused_kernel": float((idx == ours_idx).float().mean().item()),
        "symbols_ok": float((idx == syms[last]).float().mean().item()),
    }), flush=True)


if __name__ == "__main__":
    main()
