"""Preamble search over all candidate offsets (row f3 as written): lorab200_preamble_search_dev on one long
stream, stride N/4 and N/8, against what the reference does with the same samples -- its sequential hunt through
the same stream on one thread (the unmodified block in the ministream harness).  One JSON line per case."""
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

import loopback  # noqa: E402
import lora_sdr_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sf", type=int, nargs="+", default=[7, 10, 12])
ap.add_argument("--packets", type=int, default=24)
ap.add_argument("--repeat", type=int, default=8)
a = ap.parse_args()
m.build()
rng = np.random.default_rng(3)
for sf in a.sf:
    n = 1 << sf
    parts = []
    for k in range(a.packets):
        gap = int(rng.integers(n, 6 * n)) | 1
        parts.append(((rng.standard_normal(gap) + 1j * rng.standard_normal(gap)) * 0.05).astype(np.complex64))
        pl = bytes(rng.integers(0, 256, 16, dtype=np.uint8))
        parts.append(loopback.transmit(sf, 4, [pl], padding=64, sigma=0.05, seed=5 * sf + k))
    one = np.concatenate(parts)
    x = np.tile(one, a.repeat)
    ref = loopback.receive(sf, 4, one, use_b200=False, debug_ports=False, mtu=256, thresh=-3.0)
    d = m.Demod(sf)
    xd = torch.from_numpy(x).cuda()
    for div in (4, 8):
        stride = n // div
        t_end = time.perf_counter() + 0.05
        while time.perf_counter() < t_end:                       # clocks up
            rec, nf = d.preamble_search(xd, stride=stride, thresh=-3.0, min_windows=4 * div, max_found=4096)
        times = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            found = torch.zeros((4096, m.PREAMBLE_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
            nfd = torch.zeros(1, dtype=torch.int32, device="cuda")
            s = torch.cuda.current_stream().cuda_stream
            e0.record()
            st = d.L.lorab200_preamble_search_dev(d.ctx, xd.data_ptr(), xd.numel(), stride, -3.0, 4 * div,
                                                  found.data_ptr(), nfd.data_ptr(), 4096, s)
            e1.record()
            torch.cuda.synchronize()
            assert st == 0
            times.append(e0.elapsed_time(e1))
        ms = min(times)
        n_win = (x.size - n) // stride + 1
        print(json.dumps({
            "sf": sf, "stride": f"N/{div}", "samples": int(x.size), "windows": int(n_win), "preambles_found": int(nf),
            "packets_in_stream": a.packets * a.repeat, "gpu_ms": round(ms, 4),
            "Gsamples_per_s": round(x.size / ms / 1e6, 2), "Mwindows_per_s": round(n_win / ms / 1e3, 1),
            "kernel": d.last_kernel,
            "reference_hunt_one_thread_Msamples_per_s": round(one.size / ref["seconds"] / 1e6, 1),
        }), flush=True)
    d.close()
