"""Per-call latency of the C-ABI entry points for the tiny batches a block's work() passes in the
sequential receiver phases (1-2 windows per call).  Prints microseconds per call."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import lora_sdr_b200 as m  # noqa: E402

m.build()
rng = np.random.default_rng(0)
for sf in (7, 10, 12):
    n = 1 << sf
    d = m.Demod(sf)
    iq = (rng.standard_normal(4 * n) + 1j * rng.standard_normal(4 * n)).astype(np.complex64)
    dev = torch.from_numpy(iq).cuda()
    out = np.empty(4, m.RESULT_DTYPE)
    row = {}
    for name, nw, fine in (("host nw=1", 1, None), ("host nw=1 fine0", 1, (5, 0.0)), ("host nw=1 walk", 1, (5, 3.3)),
                           ("host nw=2 walk", 2, (5, 3.3)), ("host nw=4", 4, None)):
        for _ in range(50):
            d.demod_host(iq[: nw * n], nw, fine=fine, out=out[:nw])
        reps = 1000
        t0 = time.perf_counter()
        for _ in range(reps):
            d.demod_host(iq[: nw * n], nw, fine=fine, out=out[:nw])
        row[name] = (time.perf_counter() - t0) / reps * 1e6
    for name, fine in (("dev nw=1 + sync", None), ("dev nw=1 walk + sync", (5, 3.3))):
        for _ in range(50):
            d.demod_dev(dev, 1, fine=fine)
        torch.cuda.synchronize()
        reps = 1000
        t0 = time.perf_counter()
        for _ in range(reps):
            d.demod_dev(dev, 1, fine=fine)
            torch.cuda.synchronize()
        row[name] = (time.perf_counter() - t0) / reps * 1e6
    print(f"SF{sf}: " + "  ".join(f"{k}: {v:6.1f} us" for k, v in row.items()), flush=True)
    d.close()
