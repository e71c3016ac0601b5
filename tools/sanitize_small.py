"""Small fused-kernel workload for compute-sanitizer (memcheck / racecheck / synccheck):
every SF, dense + strided + constant and walking fine-tune, a partial last tile.
Usage: compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import lora_sdr_b200 as m  # noqa: E402

m.build()
rng = np.random.default_rng(1)
for sf in range(7, 13):
    n = 1 << sf
    nw = 2 * (4096 // n if n <= 1024 else 2) + 3
    nw = max(nw, 7)
    d = m.Demod(sf)
    host = (rng.standard_normal((nw + 1) * n) + 1j * rng.standard_normal((nw + 1) * n)).astype(np.complex64)
    iq = torch.from_numpy(host).cuda()
    for kw in (dict(), dict(stride=n + 2), dict(fine=(17, 0.0)), dict(fine=(17, 3.3)), dict(fine=(99, -41.7))):
        out, *_ = d.demod_dev(iq, nw, **kw)
        torch.cuda.synchronize()
        assert d.last_kernel.startswith("fused"), (sf, kw, d.last_kernel)
    d.close()
    print("sf", sf, "ok", flush=True)
