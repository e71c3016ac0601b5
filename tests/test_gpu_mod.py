"""the batched modulator (lorab200_mod_dev) against the reference's own
LoRaMod.cpp + ChirpGenerator.hpp, run unmodified through the ministream harness.

  REFERENCE mode  literal single-precision recurrences: every sample within 2 float ulp of the
                  reference's (the phases are bit-identical; polar() goes through double cos/sin
                  where the reference calls glibc cosf/sinf), the zero padding exact;
  IDEAL mode      closed-form chirps: equal to the exact chirp law to float rounding, decoded by
                  the UNMODIFIED reference receiver (demod + decoder) back to the payload bytes.
"""
import numpy as np
import pytest

import loopback

pytestmark = pytest.mark.gpu

needs_chain = pytest.mark.skipif(not loopback.available(), reason="oracle/_ref/libloraloop.so not built")


def _ideal(sf, syms, sync, padding, ampl):
    """The exact chirp law (ovs = 1): phases are integer multiples of pi/N."""
    n = 1 << sf
    out, m0 = [], 0

    def chirp(s, count, down):
        nonlocal m0
        i = np.arange(count, dtype=np.int64)
        inc = (i + 1) * (2 * s - n) + (i + 1) * (i + 2)
        m = (m0 + (-inc if down else inc)) % (2 * n)
        m0 = int(m[-1])
        return ampl * np.exp(1j * np.pi * m / n)

    for _ in range(10):
        out.append(chirp(0, n, False))
    out.append(chirp((sync >> 4) * 8, n, False))
    out.append(chirp((sync & 15) * 8, n, False))
    out.append(chirp(0, n, True))
    out.append(chirp(0, n, True))
    out.append(chirp(0, n // 4, True))
    for s in syms:
        out.append(chirp(int(s), n, False))
    out.append(np.zeros(padding * n))
    return np.concatenate(out)


@needs_chain
@pytest.mark.parametrize("sf,ovs", [(7, 1), (8, 1), (9, 1), (10, 1), (11, 1), (12, 1), (7, 2), (9, 4)])
def test_reference_mode_matches_the_reference_modulator(lib, cuda_ok, sf, ovs):
    import torch
    rng = np.random.default_rng(sf * 10 + ovs)
    ampl, sync, padding = 0.3, 0x12, 3
    packets = [loopback.encode(sf, 4, bytes(rng.integers(0, 256, 13, dtype=np.uint8))) for _ in range(5)]
    nsy = len(packets[0])
    assert all(len(p) == nsy for p in packets)
    d = lib.Demod(sf)
    syms = torch.from_numpy(np.stack(packets).astype(np.int16)).cuda()
    out = d.mod_dev(syms, sync=sync, padding=padding, ampl=ampl, ovs=ovs, mode=lib.MOD_REFERENCE).cpu().numpy()
    assert d.last_kernel == "mod_reference"
    exact = total = 0
    for p, sy in enumerate(packets):
        ref = loopback.modulate(sf, sy, padding=padding, sync=sync, ampl=ampl, ovs=ovs)
        assert ref.size == out.shape[1], "frame length differs from LoRaMod's"
        got = out[p]
        nz = ref.size - padding * (1 << sf) * ovs
        assert not ref[nz:].any() and not got[nz:].any(), "padding must be exact zeros"
        err = np.abs(got[:nz].view(np.float32) - ref[:nz].view(np.float32))
        # 2 ulp of a float in [0.25, 0.5) (|sample| <= ampl = 0.3): 2 * 2^-25
        assert err.max() <= 2 * 2.0 ** -25, f"packet {p}: {err.max()}"
        exact += int(np.sum(err == 0))
        total += err.size
    assert exact / total > 0.98, f"only {exact / total:.4f} of the components are bit-identical"
    d.close()


@needs_chain
@pytest.mark.parametrize("sf", [7, 9, 10, 12])
def test_ideal_mode_is_the_exact_chirp_law_and_the_reference_receiver_decodes_it(lib, cuda_ok, sf):
    import torch
    rng = np.random.default_rng(sf)
    ampl, sync, padding = 1.0, 0x12, 260
    payloads = [bytes(rng.integers(0, 256, 24, dtype=np.uint8)) for _ in range(4)]
    packets = [loopback.encode(sf, 4, p) for p in payloads]
    d = lib.Demod(sf)
    syms = torch.from_numpy(np.stack(packets).astype(np.int16)).cuda()
    out = d.mod_dev(syms, sync=sync, padding=padding, ampl=ampl, mode=lib.MOD_IDEAL).cpu().numpy()
    assert d.last_kernel == "mod_ideal"
    for p, sy in enumerate(packets):
        exp = _ideal(sf, sy, sync, padding, ampl)
        assert exp.size == out.shape[1]
        assert np.max(np.abs(out[p] - exp)) < 2e-7
    # the unmodified reference demod + decoder recover every payload from the GPU's frames
    rx = loopback.receive(sf, 4, out.reshape(-1), use_b200=False, debug_ports=True, mtu=256)
    assert rx["decoded"] == payloads
    # and so does our block shell
    rx2 = loopback.receive(sf, 4, out.reshape(-1), use_b200=True, debug_ports=False, mtu=256)
    assert rx2["decoded"] == payloads
    d.close()


def test_modulator_argument_checks(lib, cuda_ok):
    import torch
    d = lib.Demod(7)
    syms = torch.zeros((1, 8), dtype=torch.int16, device="cuda")
    with pytest.raises(RuntimeError):
        d.mod_dev(syms, ovs=2, mode=lib.MOD_IDEAL)      # closed form is ovs == 1 only
    with pytest.raises(RuntimeError):
        d.mod_dev(syms, ovs=0)
    assert lib.load_library().lorab200_mod_samples(7, 1, 8, 2) == (14 + 8 + 2) * 128 + 32
    d.close()
