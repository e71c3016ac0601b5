"""lorab200_fine_advance (closed-form walk of the fine-tune index, lora-sdr_b200/csrc/fine_walk.cpp)
against the sequential recurrence of LoRaDemod.cpp:160-162 as the oracle restates it
(lo_fine_next), for friendly and for deliberately awkward steps: integers, steps whose
fractional part is so small that float rounding carries in the upper binades only,
|step| < 1 (the index parks), negative steps (the index climbs and wraps down)."""
import numpy as np
import pytest

STEPS = [0.0, 1.0, -1.0, 3.0, 64.0, -17.0,            # integers
         1.28, -3.7, 55.001, 127.99, -0.015625 * 128,  # generic
         3.001, 3.0001, 2.00001, 7.000001, 100.0000076, # tiny fractional part: carry threshold inside the table
         -3.001, -2.00001, -0.999999, 0.999999,        # just inside an integer
         0.5, 0.25, 1e-3, -0.5, -1e-3,                 # |d| < 1: fixed points / parking
         511.5, -511.75, 4095.33, 65535.9]             # large steps


def _seq(oracle, x, d, length, n):
    for _ in range(n):
        x = oracle.fine_next(x, d, length)
    return x


@pytest.mark.parametrize("sf", [7, 10, 12])
def test_closed_form_equals_sequential(lib, oracle, sf):
    n = 1 << sf
    length = n * 128
    rng = np.random.default_rng(sf)
    starts = [0, 1, 2, 5, length - 1, length - 2, length // 2, n, 127, 128, 129] + \
        [int(v) for v in rng.integers(0, length, size=12)]
    counts = [1, 2, 3, n // 2, n - 1, n, n + 1, 3 * n]
    for d in STEPS:
        if abs(d) >= length:
            continue
        d32 = float(np.float32(d))
        for x0 in starts:
            # walk sequentially once, record checkpoints
            want = {}
            x, done = x0, 0
            for c in sorted(counts):
                x = _seq(oracle, x, d32, length, c - done)
                done = c
                want[c] = x
            for c in counts:
                got = lib.fine_advance(sf, x0, d32, c)
                assert got == want[c], (sf, d, x0, c, got, want[c])


def test_window_starts_of_a_long_batch(lib, oracle):
    """What the library uses it for: the index at the start of each of many windows."""
    sf, n = 9, 512
    length = n * 128
    for d in (1.28, 3.0001, -7.3, 0.3):
        d32 = float(np.float32(d))
        x_seq = x_cf = 5
        for w in range(200):
            x_seq = _seq(oracle, x_seq, d32, length, n)
            x_cf = lib.fine_advance(sf, x_cf, d32, n)
            assert x_cf == x_seq, (d, w)
        # and one big jump
        assert lib.fine_advance(sf, 5, d32, 200 * n) == x_seq


def test_random_steps_fuzz(lib, oracle):
    rng = np.random.default_rng(2024)
    sf, n = 8, 256
    length = n * 128
    for _ in range(300):
        mag = 10 ** rng.uniform(-4, 3.5)
        d32 = float(np.float32(mag * rng.choice([-1, 1])))
        x0 = int(rng.integers(0, length))
        c = int(rng.integers(1, 3 * n))
        assert lib.fine_advance(sf, x0, d32, c) == _seq(oracle, x0, d32, length, c), (d32, x0, c)


def test_bad_arguments(lib):
    with pytest.raises(lib.Lorab200Error):
        lib.fine_advance(7, -1, 1.0, 10)
    with pytest.raises(lib.Lorab200Error):
        lib.fine_advance(7, 0, 1e9, 10)
    with pytest.raises(lib.Lorab200Error):
        lib.fine_advance(6, 0, 1.0, 10)


# ---- fine_orbit.h: the walk as segments + one period (what the fused kernel evaluates in O(1)) ----------

def _orbit_lib():
    import ctypes as C
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    subprocess.run(["make", "-C", str(root / "oracle"), "libcodec_host.so"], check=True, capture_output=True)
    L = C.CDLL(str(root / "oracle" / "libcodec_host.so"))
    L.fine_orbit_host.argtypes = [C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _orbit(L, length, d32, x0, q):
    import ctypes as C
    q = np.ascontiguousarray(q, np.uint64)
    out = np.zeros(q.size, np.int32)
    info = np.zeros(3, np.int64)
    n_seg = L.fine_orbit_host(length, C.c_float(d32), x0, q.ctypes.data_as(C.c_void_p), q.size,
                              out.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p))
    return n_seg, out, info


@pytest.mark.parametrize("sf", [7, 9, 12])
def test_orbit_equals_sequential_walk(lib, oracle, sf):
    """Every sample of a walk a few wraps long (SF7/9; bounded so the Python walk stays cheap), and
    random far-away samples against the closed-form advance (itself tested against the walk above),
    for every step in STEPS that has an orbit."""
    L = _orbit_lib()
    n = 1 << sf
    length = n * 128
    rng = np.random.default_rng(sf)
    have = 0
    for d in STEPS + [13.37, 2.5, -2.5, 14.0, 127.0]:
        if d == 0.0 or abs(d) >= length:
            continue
        d32 = float(np.float32(d))
        for x0 in [0, 5, length - 1] + [int(v) for v in rng.integers(0, length, size=2)]:
            total = int(min(40000, max(3 * n, 3.3 * length / max(abs(d), 1.0)))) if sf < 12 else 2 * n
            q = np.arange(total, dtype=np.uint64)
            n_seg, got, info = _orbit(L, length, d32, x0, q)
            if n_seg == 0:
                continue
            have += 1
            x, want = x0, np.empty(total, np.int32)
            for i in range(total):
                want[i] = x
                x = oracle.fine_next(x, d32, length)
            assert np.array_equal(got, want), (sf, d, x0, int(np.argmax(got != want)), info)
            far = np.concatenate([rng.integers(0, 40 * length, size=24), [10**7 + 3, 2**33 + 12345]]).astype(np.uint64)
            _, g, _ = _orbit(L, length, d32, x0, far)
            for k, big in enumerate(far):
                assert int(g[k]) == lib.fine_advance(sf, x0, d32, int(big)), (sf, d, x0, int(big))
    assert have >= 40


def test_orbit_refuses_what_it_cannot_describe(lib):
    L = _orbit_lib()
    length = 128 * 128
    for d in (3.0001, 2.00001, 7.000001):             # finite carry threshold inside the table
        n_seg, _, _ = _orbit(L, length, float(np.float32(d)), 77, np.zeros(1, np.uint64))
        assert n_seg == 0, d
    for d in (-0.5, -1e-3):                           # D + 1 == 0: the index never moves (caller's constant path)
        n_seg, _, _ = _orbit(L, length, float(np.float32(d)), 77, np.zeros(1, np.uint64))
        assert n_seg == 0, d
