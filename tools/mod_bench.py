"""Modulator throughput (row f2): lorab200_mod_dev in both modes vs the reference LoRaMod block
(CPU, one thread, through the ministream harness).  One JSON line per SF.

IDEAL mode is a pure store kernel: algorithmic bytes = 8 B per sample written (+2 B per symbol
read), roofline = measured HBM copy peak; REFERENCE mode is bounded by the sequential recurrence
(one thread per packet) and only scales with the number of packets in the batch."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import loopback  # noqa: E402
import lora_sdr_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sf", type=int, nargs="+", default=[7, 10, 12])
ap.add_argument("--packets", type=int, default=4096)
ap.add_argument("--syms", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
m.build()
peak, src = bench.peaks()
rng = np.random.default_rng(3)
for sf in a.sf:
    n = 1 << sf
    npk = max(64, a.packets >> max(0, sf - 7))          # ~the same bytes per step at every SF
    d = m.Demod(sf)
    syms_h = rng.integers(0, n, (npk, a.syms), dtype=np.int64).astype(np.int16)
    syms = torch.from_numpy(syms_h).cuda()
    res = {}
    for mode, name in ((m.MOD_IDEAL, "ideal"), (m.MOD_REFERENCE, "reference_phase")):
        out = d.mod_dev(syms, padding=2, ampl=0.3, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = a.steps if mode == m.MOD_IDEAL else max(2, a.steps // 5)
        e0.record()
        for _ in range(steps):
            d.mod_dev(syms, padding=2, ampl=0.3, mode=mode, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        samples = out.numel()
        res[name] = {"ms_per_step": ms, "Gsamples_per_s": samples / ms / 1e6, "GBps_written": samples * 8 / ms / 1e6,
                     "frac_of_hbm_peak": samples * 8 / ms / 1e6 / peak}
    # reference LoRaMod on one CPU thread (a few packets)
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 1.0:
        ref = loopback.modulate(sf, syms_h[k % npk].astype(np.uint16), padding=2, ampl=0.3)
        k += 1
    cpu = ref.size * k / (time.perf_counter() - t0) / 1e9
    print(json.dumps({"sf": sf, "packets_per_step": npk, "symbols_per_packet": a.syms, "samples_per_step": int(samples),
                      "hbm_peak_GBps": peak, "peak_source": src, **res,
                      "reference_cpu_1thread_Gsamples_per_s": cpu}), flush=True)
    d.close()
