"""Receiver throughput (row f3): lorab200_rx_dev over many channels in one launch vs the
unmodified reference demod block (one thread per channel, here: one channel timed on one thread
through the ministream harness).  One JSON line per (SF, threshold)."""
import argparse
import json
import sys
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
ap.add_argument("--channels", type=int, default=256)
ap.add_argument("--packets", type=int, default=6)
ap.add_argument("--mtu", type=int, default=128)
a = ap.parse_args()
m.build()
rng = np.random.default_rng(11)
for sf in a.sf:
    pk = [bytes(rng.integers(0, 256, 24, dtype=np.uint8)) for _ in range(a.packets)]
    one = loopback.transmit(sf, 4, pk, padding=a.mtu, sigma=0.05, seed=sf)
    d = m.Demod(sf)
    buf = torch.from_numpy(one).cuda().repeat(a.channels, 1).contiguous()
    for thresh in (-30.0, -5.0):
        ref = loopback.receive(sf, 4, one, use_b200=False, debug_ports=True, mtu=a.mtu, thresh=thresh)
        # A launch with one busy warp per CTA is a light load: after the CPU-side reference run above the GPU
        # has idled and its clocks take tens of milliseconds to come back (the same binary on the same data
        # measured anywhere between 3.8 and 11 ms per call when timed cold).  Warm up for >= 60 ms of back to
        # back calls, then take the best of several timed calls.
        import time
        t_end = time.perf_counter() + 0.06
        while True:
            r = d.rx_dev(buf, thresh=thresh, mtu=a.mtu, max_packets=2 * a.packets + 4)
            torch.cuda.synchronize()
            if time.perf_counter() > t_end:
                break
        reps = 7
        times = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = d.rx_dev(buf, thresh=thresh, mtu=a.mtu, max_packets=2 * a.packets + 4)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = min(times)
        npk = r["n_packets"].cpu().numpy()
        info = r["info"][0].cpu().numpy()
        same = int(npk[0]) == len(ref["symbols"]) and all(
            np.array_equal(r["symbols"][0, p, : info[p, 0]].cpu().numpy(), ref["symbols"][p]) for p in range(int(npk[0])))
        print(json.dumps({
            "sf": sf, "thresh_db": thresh, "channels": a.channels, "samples_per_channel": int(one.size),
            "packets_per_channel": int(npk[0]), "identical_to_reference_block": bool(same and (npk == npk[0]).all()),
            "gpu_ms": ms, "gpu_ms_all": [round(t, 3) for t in times], "gpu_aggregate_Msamples_per_s": a.channels * one.size / ms / 1e3,
            "gpu_per_channel_Msamples_per_s": one.size / ms / 1e3,
            "reference_block_1thread_Msamples_per_s": one.size / ref["seconds"] / 1e6,
        }), flush=True)
    d.close()
