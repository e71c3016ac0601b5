"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): demodulated symbol indices bit-exact; FFT
magnitudes within a float32 tolerance.  The kernels run in "exact mode" (the
reference's float operand graph without FMA), so these tests demand more:
FFT bins, dechirped samples, index and f_index identical as numbers to the
oracle; power / power_avg bit-identical through the host-buffer entry point
(finalised with the host libm) and within POWER_TOL_DB through the
device-resident entry point (device log10).
"""
import numpy as np
import pytest

from util import bits_equal, noise_windows, same_values, synth_windows

pytestmark = pytest.mark.gpu

SFS = [7, 8, 9, 10, 11, 12]
POWER_TOL_DB = 1e-4   # absolute, dB; device log10f vs glibc log10f (both ~1-2 ulp of a value < 100)


def _dev(iq):
    import torch
    return torch.from_numpy(np.ascontiguousarray(iq)).cuda()


def _check_records(got, exp, exact_power, sf=None):
    """index / f_index / power: exact (power only via the host entry point).

    power_avg = 10*log10(total - max) - scale, where `total` is a double sum of N
    float terms.  The reference adds them sequentially, the GPU as a tree; both are
    within N*2^-53*total of the exact sum, which matters only when total - max is
    tiny against total (noise-free chirps: the value is then rounding noise in the
    reference itself).  The tolerance below is that conditioning bound; for windows
    with any real noise it is far below one float ulp, i.e. the values are identical.
    """
    assert np.array_equal(got["index"], exp["index"]), "symbol index mismatch"
    assert same_values(got["f_index"], exp["f_index"]), "f_index mismatch"
    if exact_power:
        assert same_values(got["power"], exp["power"])
    n = (1 << sf) if sf else 4096
    for k in ("power", "power_avg"):
        g, e = got[k].astype(np.float64), exp[k].astype(np.float64)
        fin = np.isfinite(e)
        assert np.array_equal(np.isfinite(g), fin)
        assert same_values(g[~fin], e[~fin])
        tol = np.full(e.shape, 0.0 if exact_power else POWER_TOL_DB)
        if k == "power_avg":
            ratio = 10.0 ** ((exp["power"].astype(np.float64) - e) / 10.0)   # max / (total - max)
            tol = tol + 8 * 4.343 * n * 2.2e-16 * (1.0 + ratio) + 2 * np.spacing(np.abs(e).astype(np.float32))
        err = np.abs(g[fin] - e[fin])
        assert np.all(err <= tol[fin]), (k, float(err.max()))


@pytest.mark.parametrize("flags", [1, 0], ids=["simple", "default"])
@pytest.mark.parametrize("sf", SFS)
def test_dense_batch_matches_oracle(lib, oracle, cuda_ok, sf, flags):
    """Dense batch of noisy chirps + pure-noise windows, ragged count."""
    n = 1 << sf
    nw = 203 if sf <= 9 else 37
    iq, syms = synth_windows(sf, nw, seed=0x10AA + sf, snr_db=0.0)
    iq[5:9] = noise_windows(sf, 4, seed=sf)          # noise-only windows
    iq[11] = 0                                       # all-zero window -> index 0, -inf powers
    d = lib.Demod(sf, flags=flags)
    up = d.table(lib.TABLE_UP)
    exp, _, efft, edec = oracle.demod_windows(sf, iq, nw, up, want_fft=True, want_dec=True)

    # device-resident entry point
    out, _, fft, dec = d.demod_dev(_dev(iq), nw, want_fft=True, want_dec=True)
    got = lib.results_to_numpy(out)
    _check_records(got, exp, exact_power=False, sf=sf)
    assert same_values(fft.cpu().numpy(), efft), "FFT bins differ from kissfft"
    assert same_values(dec.cpu().numpy(), edec), "dechirped samples differ"
    assert d.last_kernel.startswith("simple" if flags == 1 else "fused"), d.last_kernel

    # host-buffer entry point: bit-exact dB values
    got_h, _, fft_h, dec_h = d.demod_host(iq, nw, want_fft=True, want_dec=True)
    _check_records(got_h, exp, exact_power=True, sf=sf)
    assert same_values(fft_h, efft) and same_values(dec_h, edec)
    # most windows carry a chirp at 0 dB SNR: the decoded symbol is the sent one
    keep = np.ones(nw, bool); keep[5:9] = False; keep[11] = False
    assert np.mean(got["index"][keep] == syms[keep]) > 0.99
    assert got["index"][11] == 0 and np.isneginf(got["power"][11])
    d.close()


@pytest.mark.parametrize("flags", [1, 0], ids=["simple", "default"])
@pytest.mark.parametrize("sf", [7, 10, 12])
def test_stride_offset_and_down_table(lib, oracle, cuda_ok, sf, flags):
    """Windows at a stride != N starting at an odd sample (8-byte aligned only), DOWN table."""
    n = 1 << sf
    nw = 19
    stride = n + 3
    rng = np.random.default_rng(sf)
    buf = (rng.standard_normal(stride * nw + n + 1) + 1j * rng.standard_normal(stride * nw + n + 1)).astype(np.complex64)
    d = lib.Demod(sf, flags=flags)
    down = d.table(lib.TABLE_DOWN)
    exp, _, efft, _ = oracle.demod_windows(sf, buf[1:], nw, down, stride=stride, want_fft=True)
    out, _, fft, _ = d.demod_dev(_dev(buf)[1:], nw, stride=stride, table=lib.TABLE_DOWN, want_fft=True)
    _check_records(lib.results_to_numpy(out), exp, exact_power=False, sf=sf)
    assert same_values(fft.cpu().numpy(), efft)
    got_h, _, _, _ = d.demod_host(buf[1:], nw, stride=stride, table=lib.TABLE_DOWN)
    _check_records(got_h, exp, exact_power=True, sf=sf)
    d.close()


@pytest.mark.parametrize("stride_of_n", [lambda n: n + 2, lambda n: n // 2, lambda n: 2, lambda n: 3 * n],
                         ids=["n+2", "n/2", "2", "3n"])
@pytest.mark.parametrize("sf", [7, 8, 10, 12])
def test_fused_kernel_strided_and_overlapping_windows(lib, oracle, cuda_ok, sf, stride_of_n):
    """Even sample strides (16-byte aligned windows) stay on the TMA path: sliding / overlapping /
    gapped windows, e.g. a preamble search at every second sample."""
    n = 1 << sf
    stride = stride_of_n(n)
    nw = 45 if sf <= 8 else 13
    rng = np.random.default_rng(sf + stride)
    total = (nw - 1) * stride + n
    buf = (rng.standard_normal(total) + 1j * rng.standard_normal(total)).astype(np.complex64)
    d = lib.Demod(sf)
    up = d.table(lib.TABLE_UP)
    exp, _, efft, edec = oracle.demod_windows(sf, buf, nw, up, stride=stride, want_fft=True, want_dec=True)
    out, _, fft, dec = d.demod_dev(_dev(buf), nw, stride=stride, want_fft=True, want_dec=True)
    assert d.last_kernel.startswith("fused"), d.last_kernel
    _check_records(lib.results_to_numpy(out), exp, exact_power=False, sf=sf)
    assert same_values(fft.cpu().numpy(), efft) and same_values(dec.cpu().numpy(), edec)
    out2, _, _, _ = d.demod_dev(_dev(buf), nw, stride=stride)          # the non-debug variant
    _check_records(lib.results_to_numpy(out2), exp, exact_power=False, sf=sf)
    got_h, _, _, _ = d.demod_host(buf, nw, stride=stride)
    _check_records(got_h, exp, exact_power=True, sf=sf)
    d.close()


@pytest.mark.parametrize("flags", [1, 0], ids=["simple", "default"])
@pytest.mark.parametrize("step", [0.0, 1.28, -3.7, 55.001, 3.001])
@pytest.mark.parametrize("sf", [7, 9, 12])
def test_fine_tune_path(lib, oracle, cuda_ok, sf, step, flags):
    """LoRaDemod.cpp:157-166 form: in*chirp*fine[idx] with the int/float index recurrence."""
    n = 1 << sf
    nw = 23
    iq, _ = synth_windows(sf, nw, seed=77 + sf, snr_db=-5.0)
    d = lib.Demod(sf, flags=flags)
    up, fine = d.table(lib.TABLE_UP), d.table(lib.TABLE_FINE)
    idx0 = 5 if step >= 0 else n * 128 - 3
    exp, end_exp, efft, edec = oracle.demod_windows(sf, iq, nw, up, fine=fine, fine_idx=idx0,
                                                    fine_step=step, want_fft=True, want_dec=True)
    out, end, fft, dec = d.demod_dev(_dev(iq), nw, fine=(idx0, step), want_fft=True, want_dec=True)
    assert end == end_exp
    _check_records(lib.results_to_numpy(out), exp, exact_power=False, sf=sf)
    assert same_values(dec.cpu().numpy(), edec) and same_values(fft.cpu().numpy(), efft)
    got_h, end_h, _, _ = d.demod_host(iq, nw, fine=(idx0, step))
    assert end_h == end_exp
    _check_records(got_h, exp, exact_power=True, sf=sf)
    d.close()


@pytest.mark.parametrize("step", [1.28, -3.7, 55.001, 0.4, 63.99, -64.0, 17.0, 3.001953125, -0.015625])
@pytest.mark.parametrize("sf", [7, 8, 9, 10, 11, 12])
def test_fused_kernel_walking_fine_index(lib, oracle, cuda_ok, sf, step):
    """The steady-state data path of the reference (LoRaDemod.cpp:157-166 after sync: _finefreqError
    != 0, the index walks once per sample) stays on the fused kernel: per-chunk index states from
    the on-device pointer-jumping prefix, literal recurrence inside a chunk.  Batches end in a
    partial tile; two consecutive calls continue the walk (state in/out)."""
    n = 1 << sf
    nw = {7: 1031, 8: 517, 9: 259, 10: 131, 11: 67, 12: 35}[sf]
    iq, _ = synth_windows(sf, nw, seed=177 + sf, snr_db=-5.0)
    d = lib.Demod(sf)
    up, fine = d.table(lib.TABLE_UP), d.table(lib.TABLE_FINE)
    idx0 = (n * 128 * 5) // 7
    exp, end_exp, _, _ = oracle.demod_windows(sf, iq, nw, up, fine=fine, fine_idx=idx0, fine_step=step)
    out, end, _, _ = d.demod_dev(_dev(iq), nw, fine=(idx0, step))
    assert d.last_kernel.startswith("fused"), d.last_kernel
    assert end == end_exp
    _check_records(lib.results_to_numpy(out), exp, exact_power=False, sf=sf)
    # split into two calls: the returned index feeds the second call
    k = nw // 3
    o1, mid, _, _ = d.demod_dev(_dev(iq[:k]), k, fine=(idx0, step))
    o2, end2, _, _ = d.demod_dev(_dev(iq[k:]), nw - k, fine=(mid, step))
    assert end2 == end_exp
    got = np.concatenate([lib.results_to_numpy(o1), lib.results_to_numpy(o2)])
    _check_records(got, exp, exact_power=False, sf=sf)
    got_h, end_h, _, _ = d.demod_host(iq, nw, fine=(idx0, step))
    assert end_h == end_exp
    _check_records(got_h, exp, exact_power=True, sf=sf)
    d.close()


@pytest.mark.parametrize("sf,nseg,wps", [(7, 5, 12), (7, 3, 7), (8, 4, 6), (9, 6, 3), (10, 32, 8), (12, 3, 2)])
def test_multi_segment_launch(lib, oracle, cuda_ok, sf, nseg, wps):
    """lorab200_demod_multi_dev: one launch over independent per-channel device buffers; record
    c*wps + w is window w of segment c.  (7, 3, 7): wps not a multiple of the SF7 tile -> simple kernel."""
    import torch
    n = 1 << sf
    d = lib.Demod(sf)
    up = d.table(lib.TABLE_UP)
    segs, exps = [], []
    for c in range(nseg):
        iq, _ = synth_windows(sf, wps, seed=1000 * sf + c, snr_db=-3.0)
        exps.append(oracle.demod_windows(sf, iq, wps, up)[0])
        segs.append(torch.from_numpy(iq).cuda())
    out = d.demod_multi_dev(segs, wps)
    torch.cuda.synchronize()
    tile = max(1, 512 // n)
    assert d.last_kernel.startswith("fused_multi" if wps % tile == 0 else "simple_multi"), d.last_kernel
    got = lib.results_to_numpy(out)
    _check_records(got, np.concatenate(exps), exact_power=False, sf=sf)
    # strided windows inside every segment (overlapping by half a window)
    if wps > 1:
        nw2 = 2 * wps - 1
        exp2 = [oracle.demod_windows(sf, t.cpu().numpy().reshape(-1), nw2, up, stride=n // 2)[0] for t in segs]
        out2 = d.demod_multi_dev(segs, nw2, stride=n // 2)
        torch.cuda.synchronize()
        _check_records(lib.results_to_numpy(out2), np.concatenate(exp2), exact_power=False, sf=sf)
    d.close()


@pytest.mark.parametrize("sf", [7, 8, 9, 10, 11, 12])
def test_host_call_small_batches_zero_copy(lib, oracle, cuda_ok, sf):
    """lorab200_demod_host with 1..16 windows (what a block's work() passes in the sequential
    receiver phases) takes the staged zero-copy path: the records must be bit-identical to the
    oracle's, dense and strided, with and without a walking fine-tune index, call after call."""
    n = 1 << sf
    d = lib.Demod(sf)
    up, fine = d.table(lib.TABLE_UP), d.table(lib.TABLE_FINE)
    iq, _ = synth_windows(sf, 24, seed=31 + sf, snr_db=-6.0)
    flat = iq.reshape(-1)
    for nw in (1, 2, 3, 7, 16):
        if nw * n * 8 > 256 * 1024:
            continue
        exp = oracle.demod_windows(sf, flat, nw, up)[0]
        got, _, _, _ = d.demod_host(flat[: nw * n], nw)
        assert got.tobytes() == exp.tobytes(), f"dense nw={nw}"
        if nw > 1:
            stride = n // 2 + 2
            exp = oracle.demod_windows(sf, flat, nw, up, stride=stride)[0]
            got, _, _, _ = d.demod_host(flat, nw, stride=stride)
            assert got.tobytes() == exp.tobytes(), f"strided nw={nw}"
        exp, end_exp, _, _ = oracle.demod_windows(sf, flat, nw, up, fine=fine, fine_idx=77, fine_step=-9.25)
        got, end, _, _ = d.demod_host(flat[: nw * n], nw, fine=(77, -9.25))
        assert end == end_exp and got.tobytes() == exp.tobytes(), f"walking fine nw={nw}"
    d.close()


def test_detector_kat_through_gpu(lib, oracle, cuda_ok):
    """TestDetector.cpp:9-35 through the CUDA path: SF10, all 1024 symbols,
    genChirp down-chirp as the dechirp table, genChirp up-chirp(sym) as input:
    index == sym and power > -10 dB."""
    sf, n = 10, 1024
    down, _ = lib.gen_chirp(n, 1, n, 0.0, True, 1.0, 0.0)
    chirps = np.empty((n, n), np.complex64)
    for sym in range(n):
        chirps[sym], _ = lib.gen_chirp(n, 1, n, np.float32(2 * np.pi * sym) / np.float32(n), False, 1.0,
                                       np.float32(np.pi / 4))
    d = lib.Demod(sf)
    d.set_table(lib.TABLE_DOWN, down)
    got, _, _, _ = d.demod_host(chirps, n, table=lib.TABLE_DOWN)
    assert np.array_equal(got["index"], np.arange(n))
    assert np.all(got["power"] > -10.0)
    exp, _, _, _ = oracle.demod_windows(sf, chirps, n, down)
    _check_records(got, exp, exact_power=True, sf=sf)
    d.close()


def test_edge_cases(lib, cuda_ok):
    import torch
    d = lib.Demod(7)
    # zero windows is a no-op
    out, _, _, _ = d.demod_dev(torch.zeros(128, dtype=torch.complex64, device="cuda"), 0)
    assert out.shape[0] == 0
    # invalid table id
    with pytest.raises(lib.Lorab200Error):
        d.demod_dev(torch.zeros(128, dtype=torch.complex64, device="cuda"), 1, table=2)
    # bad sf
    with pytest.raises(lib.Lorab200Error):
        lib.Demod(6)
    with pytest.raises(lib.Lorab200Error):
        lib.Demod(13)
    # single window
    x = torch.zeros(128, dtype=torch.complex64, device="cuda")
    x[3] = 1.0
    out, _, _, _ = d.demod_dev(x, 1)
    assert lib.results_to_numpy(out).shape == (1,)
    d.close()
