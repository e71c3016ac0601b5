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

# the receiver kernels (speculative batches and the one-window cross-check) on a few short streams:
# frames from the device encoder + modulator, a little noise, odd lead-ins
for sf in (7, 9, 10, 12):
    for flags in (0, 1):
        d = m.Demod(sf, flags=flags)
        payloads = torch.from_numpy(rng.integers(0, 256, (3, 9), dtype=np.uint8)).cuda()
        syms, nsy = d.encode_dev(payloads, rdd=4, crc=True)
        iq = d.mod_dev(syms, padding=40, ampl=1.0, mode=m.MOD_IDEAL)
        g = torch.Generator("cuda").manual_seed(sf)
        iq = iq + 0.05 * torch.view_as_complex(torch.randn((*iq.shape, 2), device=iq.device, generator=g))
        lead = 0.05 * torch.view_as_complex(torch.randn((3, (5 << sf) // 4 + 1, 2), device=iq.device, generator=g))
        stream = torch.cat([lead, iq], dim=1).contiguous()
        for thresh in (-30.0, -5.0):
            r = d.rx_dev(stream, thresh=thresh, mtu=48, max_packets=4)
            torch.cuda.synchronize()
            assert int(r["n_packets"].min()) >= 1, (sf, flags, thresh)
        d.close()
    print("rx sf", sf, "ok", flush=True)

# the preamble search (sliding windows through the fused kernel + the mark / scan / emit kernels)
for sf in (7, 8, 10):
    n = 1 << sf
    d = m.Demod(sf)
    payloads = torch.from_numpy(rng.integers(0, 256, (1, 9), dtype=np.uint8)).cuda()
    syms, nsy = d.encode_dev(payloads, rdd=4, crc=True)
    iq = d.mod_dev(syms, padding=20, ampl=1.0, mode=m.MOD_IDEAL)[0]
    g = torch.Generator("cuda").manual_seed(50 + sf)
    lead = 0.05 * torch.view_as_complex(torch.randn((n + 3, 2), device=iq.device, generator=g))
    stream = torch.cat([lead, iq, lead, iq]).contiguous()
    stream = stream + 0.05 * torch.view_as_complex(torch.randn((stream.numel(), 2), device=iq.device, generator=g))
    for div in (4, 8):
        rec, nf = d.preamble_search(stream, stride=n // div, thresh=-3.0, min_windows=4 * div)
        assert nf == 2, (sf, div, nf)
    d.close()
    print("search sf", sf, "ok", flush=True)
