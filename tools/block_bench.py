"""Drop-in measurement at the BLOCK boundary: the same IQ stream through
    reference LoRaDemod.cpp block  -> reference decoder      (CPU, unmodified reference code)
    LoRaDemodB200 block shell      -> reference decoder      (our kernels behind the block API)
on the ministream scheduler (oracle/_ref/libloraloop.so).  The stream is a burst of LoRa packets
made by the reference encoder + modulator, repeated; both chains must decode the same packets.
Reports stream rate (Msamples/s) and the wall time of the scheduler loop for each demod.

    python tools/block_bench.py --sf 7 10 12 --packets 16 --repeat 8
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import loopback  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sf", type=int, nargs="+", default=[7, 10, 12])
ap.add_argument("--packets", type=int, default=16)
ap.add_argument("--payload", type=int, default=32)
ap.add_argument("--mtu", type=int, default=192, help="demod MTU and modulator padding, in symbols")
ap.add_argument("--repeat", type=int, default=8)
ap.add_argument("--cr", type=int, default=4)
ap.add_argument("--thresh", type=float, default=-30.0, help="demod squelch threshold in dB (reference default -30 = never squelched)")
a = ap.parse_args()
rng = np.random.default_rng(7)
for sf in a.sf:
    pk = [bytes(rng.integers(0, 256, a.payload, dtype=np.uint8)) for _ in range(a.packets)]
    iq1 = loopback.transmit(sf, a.cr, pk, padding=a.mtu, sigma=0.05)
    iq = np.tile(iq1, a.repeat)
    res = {}
    # best of three runs each (a run is a fresh block: context creation and slab allocation are outside the
    # timed loop, but a shared host and a GPU that has just been idle make single runs noisy)
    for name, use in (("b200", True), ("reference", False), ("b200", True), ("reference", False), ("b200", True), ("reference", False)):
        r = loopback.receive(sf, a.cr, iq, use_b200=use, debug_ports=False, mtu=a.mtu, thresh=a.thresh)
        if name not in res or r["seconds"] < res[name]["seconds"]:
            res[name] = r
    ref, got = res["reference"], res["b200"]
    ok = got["n_decoded"] == ref["n_decoded"] and got["decoded"] == ref["decoded"] and \
        all(np.array_equal(x, y) for x, y in zip(got["symbols"], ref["symbols"]))
    print(json.dumps({
        "sf": sf, "thresh_db": a.thresh, "samples": int(iq.size), "packets_sent": a.packets * a.repeat,
        "packets_decoded": {"reference": ref["n_decoded"], "b200": got["n_decoded"]}, "identical": bool(ok),
        "reference_block": {"seconds": ref["seconds"], "Msamples_per_s": iq.size / ref["seconds"] / 1e6},
        "b200_block": {"seconds": got["seconds"], "Msamples_per_s": iq.size / got["seconds"] / 1e6,
                       "demod_calls": got["calls"], "windows_submitted": got["windows"], "stream_windows": int(iq.size >> sf), "kernel_launches": got["launches"]},
        "speedup": ref["seconds"] / got["seconds"],
    }), flush=True)
