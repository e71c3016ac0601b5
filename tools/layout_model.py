#!/usr/bin/env python3
"""Executable model of the fused kernel's data movement (design aid + regression check).

For every SF 7..12 it replays, in numpy float64, exactly the per-thread register /
shared-memory choreography that lora-sdr_b200/csrc/demod_fused.cuh implements:

  pass 0   thread t loads 16 samples (E sub-groups of R0 = 2^b0 points, strided by
           N/R0 through the linear TMA stage buffer), dechirps, runs the deepest
           kissfft stages in registers, scatters F[j] to the padded exchange layout
  pass 1.. thread (lo, hi) gathers 16 values F[lo | y<<a | hi<<(a+4)], runs two
           radix-4 stages with its own twiddles, scatters / finishes

and checks (1) the result equals numpy's FFT of the dechirped window, i.e. the
index bookkeeping is right, and (2) every shared-memory access instruction is
bank-conflict free under the hardware rule (64-bit accesses: per half-warp, 16
distinct 8-byte bank pairs; 128-bit: per quarter-warp, 8 distinct 16-byte slots).

Run:  python tools/layout_model.py
"""
import sys

import numpy as np


def cfg(sf):
    n = 1 << sf
    npass = (sf + 3) // 4
    b0 = sf - 4 * (npass - 1)
    g = n // 16
    group = max(32, g)
    return dict(sf=sf, n=n, npass=npass, b0=b0, r0=1 << b0, e=16 >> b0, g=g, group=group,
                wpt=max(1, 512 // n))


def digit_rev(v, nbits):
    """reverse base-4 digits of v over nbits (nbits even)"""
    r = 0
    for _ in range(nbits // 2):
        r = (r << 2) | (v & 3)
        v >>= 2
    return r


def ilow(c, t, e):
    if c["e"] == 1:
        return t
    e0, ehi = e & 1, e >> 1
    return ehi * (2 * c["g"]) + 2 * t + e0


def x_of_c(b0, cc):
    """pass-0 register slot c (input order, stride N/R0) -> low j bits x"""
    if b0 <= 2:
        return cc
    if b0 == 3:
        return ((cc & 3) << 1) | (cc >> 2)
    return ((cc & 3) << 2) | (cc >> 2)


SF_NOW = 7


def pad(j):
    """exchange layout (FusedCfg::pad): pads keep every address = thread base + immediate.
    SF7 (two windows per half-warp) needs its own pattern and a window stride of 152."""
    if SF_NOW == 7:
        return j + (j >> 3) + 2 * (j >> 5)
    return j + (j >> 4) + (j >> 8)


def win_stride(c):
    return 152 if c["sf"] == 7 else pad(c["n"] - 1) + 1


def bfly4(a0, a1, a2, a3):
    s5 = a0 - a2
    f0 = a0 + a2
    s3 = a1 + a3
    s4 = a1 - a3
    s4r = complex(s4.imag, -s4.real)
    return f0 + s3, s5 + s4r, f0 - s3, s5 - s4r


def model_window(sf, xin, chirp, tw, accesses):
    global SF_NOW
    SF_NOW = sf
    c = cfg(sf)
    n, b0, r0, E, g = c["n"], c["b0"], c["r0"], c["e"], c["g"]
    hi_bits = sf - b0
    exch = {}
    # ---- pass 0
    for t in range(g):
        v = {}
        for e in range(E):
            for cc in range(r0):
                i = ilow(c, t, e) + (n // r0) * cc
                if E == 1 or (e & 1) == 0:   # E > 1: one LDS.128 fetches the (e0=0, e0=1) pair
                    accesses.setdefault(("p0read", e >> 1 if E > 1 else 0, cc), []).append((t, i, 16 if E > 1 else 8))
                v[(e, cc)] = xin[i] * chirp[i]
        for e in range(E):
            if b0 == 1:
                a, b = v[(e, 0)], v[(e, 1)]
                v[(e, 0)], v[(e, 1)] = a + b, a - b
            elif b0 == 2:
                v[(e, 0)], v[(e, 1)], v[(e, 2)], v[(e, 3)] = bfly4(v[(e, 0)], v[(e, 1)], v[(e, 2)], v[(e, 3)])
            elif b0 == 3:
                for a in range(4):
                    p, q = v[(e, a)], v[(e, a + 4)]
                    v[(e, a)], v[(e, a + 4)] = p + q, p - q
                for k in range(2):
                    r = [v[(e, q + 4 * k)] for q in range(4)]
                    r = [r[q] * tw[q * k * (n // 8)] for q in range(4)]
                    r = bfly4(*r)
                    for q in range(4):
                        v[(e, q + 4 * k)] = r[q]
            else:
                for qb in range(4):
                    r = bfly4(*[v[(e, qb + 4 * qa)] for qa in range(4)])
                    for qa in range(4):
                        v[(e, qb + 4 * qa)] = r[qa]
                for k in range(4):
                    r = [v[(e, q + 4 * k)] * tw[q * k * (n // 16)] for q in range(4)]
                    r = bfly4(*r)
                    for q in range(4):
                        v[(e, q + 4 * k)] = r[q]
        for e in range(E):
            jhi = digit_rev(ilow(c, t, e), hi_bits)
            for cc in range(r0):
                j = x_of_c(b0, cc) | (jhi << b0)
                exch[j] = v[(e, cc)]
                accesses.setdefault(("x0write", e, cc), []).append((t, pad(j), 8))
    # ---- passes 1..
    # Two-pass sizes: thread t has lo = t.  Three-pass sizes: physical thread tau plays
    # (lo, hi) = (tau >> 4, tau & 15) in pass 1 and (lo, m) = (tau >> 4, tau & 15) in pass 2
    # (logical pass-2 coordinate lo | m << b0), so that the second exchange only moves data
    # between the 16 threads of one lo-group, i.e. inside a warp.
    a = b0
    F = exch
    for p in range(1, c["npass"]):
        out = {}
        for tau in range(g):
            if c["npass"] == 2:
                lo, hi = tau, 0
            elif p == 1:
                lo, hi = tau >> 4, tau & 15
            else:
                lo, hi = (tau >> 4) | ((tau & 15) << b0), 0
            r = []
            for y in range(16):
                j = lo | (y << a) | (hi << (a + 4))
                accesses.setdefault((f"x{p-1}read", y), []).append((tau, pad(j), 8))
                r.append(F[j])
            m = 1 << a
            fsa = n // (4 * m)
            for qd in range(4):
                s = [r[qd * 4 + qc] * tw[qc * lo * fsa] for qc in range(4)]
                s = bfly4(*s)
                for qc in range(4):
                    r[qd * 4 + qc] = s[qc]
            mb = m * 4
            fsb = n // (4 * mb)
            for qc in range(4):
                k = lo + qc * m
                s = [r[qd * 4 + qc] * tw[qd * k * fsb] for qd in range(4)]
                s = bfly4(*s)
                for qd in range(4):
                    r[qd * 4 + qc] = s[qd]
            for y in range(16):
                j = lo | (y << a) | (hi << (a + 4))
                out[j] = r[y]
                if p + 1 < c["npass"]:
                    accesses.setdefault((f"x{p}write", y), []).append((tau, pad(j), 8))
        F = out
        a += 4
    return np.array([F[j] for j in range(n)])


def check_banks(sf, accesses):
    global SF_NOW
    SF_NOW = sf
    c = cfg(sf)
    g, wpt, n = c["g"], c["wpt"], c["n"]
    bad = []
    for key, acc in accesses.items():
        linear = key[0] == "p0read"
        wstride = n if linear else win_stride(c)
        # build per-lane element addresses for the warps of one group
        per_thread = {}
        for (t, el, nbytes) in acc:
            per_thread.setdefault(t, []).append((el, nbytes))
        # one instruction = one entry per thread
        assert all(len(v) == 1 for v in per_thread.values()), key
        lanes = []
        for w in range(wpt):
            for t in range(g):
                el, nbytes = per_thread[t][0]
                lanes.append((w * wstride + el, nbytes))
        for w0 in range(0, len(lanes), 32):
            warp = lanes[w0:w0 + 32]
            nbytes = warp[0][1]
            chunk = 16 if nbytes == 8 else 8
            for h in range(0, 32, chunk):
                part = warp[h:h + chunk]
                if nbytes == 8:
                    slots = [el % 16 for el, _ in part]
                else:
                    slots = [(el // 2) % 8 for el, _ in part]
                deg = max(slots.count(s) for s in set(slots))
                if deg > 1:
                    bad.append((key, w0 // 32, h, deg))
    return bad


def fine_pad(j):
    """FINE_WALK: word offset of fine-tune index j in a window's exchange area (demod_fused.cuh)"""
    return j + 4 * (j >> 5)


def check_fine_layout(sf):
    """FINE_WALK index parking: thread t writes words fine_pad(16t + i) as four 128-bit stores and
    reads its pass-0 samples' indices (32- or 64-bit loads).  Checks (1) the reader's address is
    thread base + immediate, (2) write then read returns index j for sample j, (3) bank conflicts
    (words, 32 banks; 128-bit: per quarter warp, 64-bit: per half warp, 32-bit: per warp)."""
    c = cfg(sf)
    n, g, wpt, e, r0 = c["n"], c["g"], c["wpt"], c["e"], c["r0"]
    fstride = n + n // 8 + (16 if sf == 8 else 0)
    assert fstride * 4 % 16 == 0
    mem = {}
    bad = []
    lanes_of_warp = [(w, t) for w in range(wpt) for t in range(g)]
    # writes
    for q in range(4):
        for w0 in range(0, len(lanes_of_warp), 32):
            warp = lanes_of_warp[w0:w0 + 32]
            addrs = []
            for (w, t) in warp:
                a = w * fstride + fine_pad(16 * t) + 4 * q
                assert a % 4 == 0
                for i in range(4):
                    assert (w, a + i) not in mem
                    mem[(w, a + i)] = 16 * t + 4 * q + i
                addrs.append(a)
            for h in range(0, len(warp), 8):
                banks = [(a // 4) % 8 for a in addrs[h:h + 8]]
                if max(banks.count(b) for b in set(banks)) > 1:
                    bad.append(("fine write", q, w0 // 32, h))
    # reads
    width = 1 if e == 1 else 2
    insts = [(0, cc) for cc in range(r0)] if e == 1 else [(eh, cc) for eh in range(e // 2) for cc in range(r0)]
    for (eh, cc) in insts:
        jc = eh * 2 * g + (n // r0) * cc
        for w0 in range(0, len(lanes_of_warp), 32):
            warp = lanes_of_warp[w0:w0 + 32]
            addrs = []
            for (w, t) in warp:
                il = t if e == 1 else 2 * t
                a = w * fstride + fine_pad(il) + fine_pad(jc)          # thread base + immediate
                assert fine_pad(il) + fine_pad(jc) == fine_pad(il + jc), (sf, t, jc)
                for i in range(width):
                    assert mem[(w, a + i)] == il + jc + i
                addrs.append(a)
            per = 32 if width == 1 else 16
            for h in range(0, len(warp), per):
                banks = [(a // width) % (32 // width) for a in addrs[h:h + per]]
                if max(banks.count(b) for b in set(banks)) > 1:
                    bad.append(("fine read", eh, cc, w0 // 32, h))
    return bad


def main():
    rng = np.random.default_rng(0)
    ok = True
    for sf in range(7, 13):
        c = cfg(sf)
        n = c["n"]
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        chirp = np.exp(1j * rng.uniform(0, 2 * np.pi, n))
        tw = np.exp(-2j * np.pi * np.arange(n) / n)
        acc = {}
        got = model_window(sf, x, chirp, tw, acc)
        ref = np.fft.fft(x * chirp)
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        bad = check_banks(sf, acc) + check_fine_layout(sf)
        print(f"SF{sf}: b0={c['b0']} npass={c['npass']} G={c['g']} E={c['e']} fft err {err:.2e}  "
              f"exchange footprint/window {win_stride(c)} el  bank conflicts: {len(bad)}")
        for b in bad[:8]:
            print("    ", b)
        ok &= err < 1e-12 and not bad
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
