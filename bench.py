#!/usr/bin/env python3
"""bench.py -- Msymbols/s of the fused LoRa demod (dechirp + FFT + argmax) on B200.

Contract (see the task text): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  A "step" is one pass of the hot path over one
batch of synthetic IQ already resident in HBM:

  workload (N=1 and per GPU for N>1, weak scaling): BASELINE.json configs[1] --
  SF=7 (128-sample windows), 65 536 dense symbol windows (64 MiB complex64) per
  step, each window a reference-recipe up-chirp at a random symbol + AWGN at
  0 dB per-sample SNR.  Other SFs / sizes: --sf / --windows (parity-test cases,
  not the headline).

  value      whole-job Msymbols/s, inputs resident in HBM, CUDA-event timed on the
             launching stream over exactly K steps replayed from one CUDA graph
             (a 64 MiB step lasts ~10-20 us: enqueueing K launches from Python
             would time the interpreter, not the GPU); max over ranks.
  e2e        same metric through the host-buffer C-ABI call (lorab200_demod_host:
             pinned host IQ -> H2D -> kernel -> D2H results -> host finalise), every
             copy inside the timed region; the last step's records are compared with
             the CPU oracle (e2e.check).
  roofline   HBM: algorithmic bytes (8N+16 per window) / average kernel time.
  sweep      the same device-timed measurement at the other SFs BASELINE.json's metric
             names (SF10, SF12 x 16 384 windows) and with the fine-tune multiply of
             LoRaDemod.cpp:157-166 (constant and walking index); `cfg4` = BASELINE
             configs[3]: 256 channels x SF10 x 1 024 windows, the channels sharded
             over the ranks, one multi-channel launch per GPU per step.
  cpu_baseline  the reference's own kissfft path (oracle/_ref, or the port) on the
             host cores, bounded sample: all threads, one thread (the reference is
             single-threaded per block) and the -O3/AVX2 courtesy build.

`--impl reference` times the reference CPU implementation instead (rank 0 only).
L2 hygiene: steps rotate over enough distinct input buffers to exceed 3x the
126 MB L2, so every step streams its IQ from HBM.
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

L2_BYTES = 126 * 1024 * 1024
METRIC = "Msymbols/s demod (dechirp+FFT+argmax)"
CFG4_CHANNELS, CFG4_SF, CFG4_WINDOWS = 256, 10, 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=int, default=7)
    ap.add_argument("--windows", type=int, default=65536)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (0 = min(K, 100))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the SF10/SF12/fine-tune/cfg4 sweep records")
    ap.add_argument("--simple", action="store_true", help="force the simple (non-fused) kernel")
    ap.add_argument("--l2-resident", action="store_true",
                    help="reuse ONE input buffer every step (fits the 126 MB L2 at the default size): the kernel's "
                         "speed with DRAM out of the picture; not the headline (the default rotates > 3x L2)")
    ap.add_argument("--channels", type=int, default=1,
                    help="BASELINE config 4 form: this many independent channel buffers of --windows windows each, "
                         "demodulated by ONE lorab200_demod_multi_dev launch per step")
    ap.add_argument("--fine-step", type=float, default=None,
                    help="demodulate with a walking fine-tune index of this step per sample "
                         "(LoRaDemod.cpp:157-166 after sync); default: the detector-only form")
    ap.add_argument("--no-input-stable", action="store_true",
                    help="do not set LORAB200_FLAG_INPUT_STABLE (first loads wait for the previous launch to drain)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(sf, windows):
    return f"SF={sf} (N={1 << sf}) batched demod, {windows} synthetic symbol windows, dense complex64"


def n_buffers(step_bytes, l2_resident=False):
    if l2_resident:
        return 1
    return min(64, max(2, int(np.ceil(3 * L2_BYTES / step_bytes)) + 1))


def make_config(args):
    """The `config` object of the JSON line: a pure function of the command line, so that the
    b200 arm and the `--impl reference` arm of one driver run print the SAME object."""
    sf, nch, wpc = args.sf, args.channels, args.windows
    nw = wpc * nch
    step_bytes = nw * (1 << sf) * 8
    nbuf = n_buffers(step_bytes, args.l2_resident)
    return {
        "workload": workload_name(sf, nw) if nch == 1 else
        f"{nch} channels x SF={sf} x {wpc} windows per channel per GPU, one multi-channel launch per step",
        "sf": sf, "windows_per_step_per_gpu": nw, "channels_per_gpu": nch,
        "l2": (f"inputs rotate over {nbuf} distinct {step_bytes >> 20} MiB buffers "
               f"({nbuf * step_bytes >> 20} MiB > 3x L2), so every step streams from HBM") if not args.l2_resident
        else f"ONE {step_bytes >> 20} MiB buffer reused every step (L2-resident, --l2-resident)",
        "timing": "K steps captured in one CUDA graph (programmatic dependent launch between steps, "
                  "LORAB200_FLAG_INPUT_STABLE), CUDA events on the launching stream around exactly those K "
                  "steps; after the barrier + synchronize the same graph runs untimed for ~4 ms directly in "
                  "front of the timed replay (a B200 that idled at the barrier runs its next ~2 ms 5 % slower)",
        "fine_step": args.fine_step,
        "sharding": "windows block-partitioned per rank, no data-path collective",
    }


def synth_on_device(sf, windows, seed, device):
    """Reference-recipe up-chirp at a random symbol per window + AWGN (0 dB/sample), built on the GPU."""
    import torch

    n = 1 << sf
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    syms = torch.randint(0, n, (windows,), generator=g, device=device)
    i = torch.arange(n, device=device, dtype=torch.float64)
    ph = -np.pi * (i + 1) + np.pi * i * (i + 1) / n          # phase of the demod's table (LoRaDemod.cpp:97-107)
    iq = torch.empty((windows, n), dtype=torch.complex64, device=device)
    blk = max(1, (1 << 22) // n)                              # bounded float64 temporaries
    for a in range(0, windows, blk):
        b = min(windows, a + blk)
        phase = ph[None, :] + (2 * np.pi / n) * syms[a:b, None].double() * i[None, :]
        phase = torch.remainder(phase, 2 * np.pi)
        sig = torch.polar(torch.ones_like(phase), phase).to(torch.complex64)
        noise = torch.randn((b - a, n, 2), generator=g, device=device, dtype=torch.float32) * np.sqrt(0.5)
        iq[a:b] = sig + torch.view_as_complex(noise)
    return iq, syms


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while a timed region runs."""

    def __init__(self, index, period_s=0.001):
        self.period_s = period_s
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = None

    def _sample(self):
        nv = self.nv
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self._names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(self.period_s)

    def start(self, background=True):
        """One sample at once, then (background=True) one per period from a thread.  NVML queries
        perturb a short timed region (measured: +0.6 .. 1.4 us per 14 us step over a 20-step region),
        so regions shorter than a few sampling periods are only bracketed: one sample before the
        region is enqueued, one after it has been synchronised."""
        if self.nv is None:
            return
        nv = self.nv
        self._names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        self._sample()
        self._stop.clear()
        self._started = True
        if not background:
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        """Call BEFORE synchronising on the timed work where possible; takes a last sample."""
        if self.th is not None:
            self._stop.set()
            self.th.join()
            self.th = None
            self._sample()
        elif getattr(self, "_started", False):
            self._sample()
        self._started = False

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- host placement ------------------------------------------------------------------------------------

def _cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(local, n_local):
    """Bind this rank (and the helper threads the library starts later) to the CPUs of its GPU's
    NUMA node, split evenly among the local ranks whose GPUs hang off the same node, BEFORE any
    pinned memory is allocated: page-locked buffers are first-touched on the local node and the
    ranks do not compete for the same cores (round 1: e2e efficiency 0.80 at 8 GPUs without it)."""
    import torch
    try:
        def node_of(i):
            p = torch.cuda.get_device_properties(i)
            dev = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            return int(Path(f"/sys/bus/pci/devices/{dev}/numa_node").read_text())
        node = node_of(local)
        if node < 0:
            return {"bound": False, "why": "numa_node unknown (-1)"}
        peers = [i for i in range(n_local) if node_of(i) == node]
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in _cpulist(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text()) if c in allowed]
        if not cpus:
            return {"bound": False, "why": f"no allowed CPU on node {node}"}
        k, m = peers.index(local), len(peers)
        per = max(1, len(cpus) // m)
        share = cpus[k * per:(k + 1) * per] or cpus
        os.sched_setaffinity(0, share)
        return {"bound": True, "numa_node": node, "cpus": len(share), "ranks_on_node": m}
    except Exception as e:  # no sysfs / no permission: run unbound and say so
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


# ---- CPU baseline ------------------------------------------------------------------------------------

def _reference_impl(o3=False):
    from oracle import pyoracle
    try:
        return pyoracle.Reference(o3=o3)
    except Exception:
        return None if o3 else pyoracle.Oracle()


def _time_cpu(impl, sf, windows_host, up, nthreads, budget_s, reps=3):
    """Best of `reps` passes over a sample sized to about budget_s / reps seconds each."""
    nw = windows_host.shape[0]
    probe = min(nw, max(64, 32 * nthreads))
    t0 = time.perf_counter()
    impl.demod_windows(sf, windows_host[:probe], probe, up, nthreads=nthreads)
    rate = probe / (time.perf_counter() - t0)
    sample = int(min(nw, max(probe, rate * budget_s / reps)))
    best, res = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        res, _, _, _ = impl.demod_windows(sf, windows_host[:sample], sample, up, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample, sample / best / 1e6, res


def cpu_baseline(sf, windows_host, cores):
    """cpu_baseline object: the reference's own CPU path on all host threads, on one thread, and the
    -O3/AVX2 courtesy build, each on a bounded sample of the step's batch."""
    impl = _reference_impl()
    up, _ = impl.tables_demod(sf)
    sample, v_all, _ = _time_cpu(impl, sf, windows_host, up, cores, budget_s=6.0)
    s1, v_one, _ = _time_cpu(impl, sf, windows_host, up, 1, budget_s=2.0, reps=2)
    out = {"value": v_all, "unit": "Msymbols/s", "cores": cores, "kind": impl.kind,
           "sample": f"first {sample} windows of the step's batch, {cores} threads, one LoRaDetector per thread, "
                     f"best of 3 passes; g++ -O2, no -march (the reference's flagless CMake)",
           "one_thread": {"value": v_one, "unit": "Msymbols/s", "cores": 1,
                          "sample": f"first {s1} windows, best of 2 passes "
                                    "(the reference runs one thread per block)"}}
    o3 = _reference_impl(o3=True)
    if o3 is not None:
        s3, v3, _ = _time_cpu(o3, sf, windows_host, up, cores, budget_s=3.0)
        s4, v4, _ = _time_cpu(o3, sf, windows_host, up, 1, budget_s=1.0, reps=2)
        out["o3_avx2"] = {"value": v3, "one_thread": v4, "unit": "Msymbols/s", "cores": cores,
                          "flags": "-O3 -march=x86-64-v3 -ffp-contract=off (courtesy build, identical results)",
                          "sample": f"first {s3} windows ({s4} for one thread)"}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.

    A step is a bounded sample of the workload: REPS passes over the step's batch (the same bytes the
    b200 arm demodulates) inside one timed region, so that thread wake-up and imbalance of a 3-7 ms
    fork/join stop dominating; the value is the best of K steps (round 1 saw 9.4 .. 25 Msym/s for the
    mean of K short steps on the shared host)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sf, nw = args.sf, args.windows * args.channels
    sys.path.insert(0, str(ROOT / "tests"))
    from util import synth_windows

    impl = _reference_impl()
    cores = len(os.sched_getaffinity(0))
    up, _ = impl.tables_demod(sf)
    probe_iq, _ = synth_windows(sf, 2048, seed=1, snr_db=0.0)
    t0 = time.perf_counter()
    impl.demod_windows(sf, probe_iq, 2048, up, nthreads=cores)
    rate = 2048 / (time.perf_counter() - t0)
    # ~90 s for the whole run; at least 4 passes' worth of windows per step when the budget allows
    total_budget = 90.0
    n_steps = max(1, args.steps + args.warmup)
    per_step = int(min(4 * nw, max(1024, rate * total_budget / n_steps)))
    base = min(per_step, nw)
    iq1, _ = synth_windows(sf, base, seed=0x10AA, snr_db=0.0)
    reps = max(1, per_step // base)
    iq = np.ascontiguousarray(np.tile(iq1, (reps, 1)))
    per_step = iq.shape[0]
    for _ in range(args.warmup):
        impl.demod_windows(sf, iq, per_step, up, nthreads=cores)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        impl.demod_windows(sf, iq, per_step, up, nthreads=cores)
        times.append(time.perf_counter() - t0)
    best, mean = min(times), float(np.mean(times))
    v = per_step / best / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msymbols/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": make_config(args),
        "cpu_baseline": {"value": v, "unit": "Msymbols/s", "cores": cores, "kind": impl.kind,
                         "sample": f"{per_step} windows per step ({reps} x the first {base} windows of the "
                                   f"{nw}-window workload), best of {args.steps} steps "
                                   f"(mean {per_step / mean / 1e6:.2f} Msym/s), {cores} threads"},
        "e2e": {"value": v, "unit": "Msymbols/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- device-resident timing ----------------------------------------------------------------------------

def time_graph(torch, dist, world, device, launch, K, W, demod, sampler=None, sustain_s=0.0):
    """W warm-up launches, then exactly K launches replayed from ONE CUDA graph, CUDA events on the
    launching stream, barrier + synchronize on both sides, max over ranks.  Returns (ms, launches,
    sustained) -- `sustained` is None unless sustain_s > 0: the same graph replayed back to back for
    about sustain_s seconds, the LAST replay timed, with the clocks seen (on B200 this kernel reaches
    the software power cap after ~70 ms of continuous running and the SM clock drops to ~1.8 GHz; the
    K-step figure is the burst number, like the copy MEASURED_PEAKS.json's hbm_gbs was measured with)."""
    stream = torch.cuda.Stream(device=device)
    with torch.cuda.stream(stream):
        for i in range(max(3, W)):
            launch(i, stream.cuda_stream)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        launches0 = demod.kernel_launches
        with torch.cuda.graph(graph, stream=stream):
            for i in range(K):
                launch(i, torch.cuda.current_stream().cuda_stream)
        launches = demod.kernel_launches - launches0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        graph.replay()                       # warm replay (graph upload); its length decides how the
        e1.record(stream)                    # timed region is sampled and pre-loaded
        stream.synchronize()
    est_ms = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.start(background=est_ms >= 20.0)
        if est_ms < 20.0:
            time.sleep(0.005)                # let the NVML query's own disturbance die out
    # Warm-up, continued: the K-step graph replayed untimed for ~4 ms, then the barrier + synchronize the
    # contract asks for, then the timed K steps at once.  Measured with tools/kbench (--idle-ms,
    # --print-reps): after an idle gap of >= 1 ms a B200 runs this kernel ~5 % slower and needs ~2 ms of
    # continuous work to get back to its steady speed, so a 20-step (0.3 ms) region timed right after
    # the interpreter's pauses (NVML queries, graph capture) would time that ramp, not the kernel.
    # ~4 ms is far below the ~70 ms after which the software power cap sets in (see `sustained`).
    # With several ranks the barrier itself is such a gap (a rank that arrives first idles until the last one
    # does: the 8-GPU line measured 15.5 us per step against 14.4 on 1-2 GPUs while the longer sweep
    # entries scaled at 0.998-1.0), so the pre-load is enqueued AFTER the barrier + synchronize, directly in
    # front of the timed graph on the same stream: the events still bracket exactly K steps.
    n_pre = 0 if est_ms >= 8.0 else max(1, min(200, int(np.ceil(4.0 / max(est_ms, 1e-3)))))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(n_pre):
            graph.replay()
        e0.record(stream)
        graph.replay()
        e1.record(stream)
    stream.synchronize()
    if sampler is not None:
        sampler.stop()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sustained = None
    if sustain_s > 0:
        reps = max(2, int(sustain_s * 1e3 / max(ms, 1e-3)))
        s2 = ClockSampler(device.index, period_s=0.01)
        s2.start()
        with torch.cuda.stream(stream):
            for _ in range(reps - 1):
                graph.replay()
            e0.record(stream)
            graph.replay()
            e1.record(stream)
        stream.synchronize()
        s2.stop()
        ts = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sustained = {"ms_per_step": float(ts.item()) / K, "replays": reps, "seconds": reps * ms * 1e-3,
                     "clocks": s2.summary()}
    return ms, int(launches), sustained


def oracle_check(m, d, sf, iq, got, n_check=256):
    """A sample of windows of `iq` (device tensor [W][N]) against the CPU oracle: indices bit-exact."""
    import torch
    from oracle import pyoracle

    nw = iq.shape[0]
    pick = np.linspace(0, nw - 1, num=min(nw, n_check), dtype=np.int64)
    sub = iq[torch.from_numpy(pick).to(iq.device)].cpu().numpy()
    exp, _, _, _ = pyoracle.Oracle().demod_windows(sf, sub, len(pick), d.table(m.TABLE_UP))
    return pick, exp, bool(np.array_equal(got["index"][pick], exp["index"]))


def sweep_entry(torch, dist, m, world, rank, device, local, sf, windows, steps, fine, peak, label):
    """One device-timed record of the sweep: same method as the headline (graph of `steps` launches,
    inputs rotating over > 3x L2), indices of a sample checked against the oracle."""
    n = 1 << sf
    d = m.Demod(sf, device=local, flags=m.FLAG_INPUT_STABLE)
    step_bytes = windows * n * 8
    nbuf = n_buffers(step_bytes)
    bufs = [synth_on_device(sf, windows, seed=0x10AA + sf + 7919 * b + 104729 * rank, device=device)[0]
            for b in range(nbuf)]
    out = torch.empty((windows, 4), dtype=torch.int32, device=device)
    d.demod_dev(bufs[0], windows, out=out, fine=fine)
    torch.cuda.synchronize()
    got = m.results_to_numpy(out)
    if fine is None:
        _, _, ok = oracle_check(m, d, sf, bufs[0], got, n_check=64)
    else:
        # with the fine-tune multiply: window w starts at table index advance(index0, step, w*N)
        from oracle import pyoracle
        orc, ok = pyoracle.Oracle(), True
        up, ftab = d.table(m.TABLE_UP), d.table(m.TABLE_FINE)
        for w in np.linspace(0, windows - 1, num=16, dtype=np.int64):
            idx = m.fine_advance(sf, fine[0], fine[1], int(w) * n)
            exp, _, _, _ = orc.demod_windows(sf, bufs[0][int(w)].cpu().numpy()[None, :], 1, up,
                                             fine=ftab, fine_idx=idx, fine_step=fine[1])
            ok = ok and int(exp["index"][0]) == int(got["index"][w])

    def launch(i, s):
        d.demod_dev(bufs[i % nbuf], windows, out=out, stream=s, fine=fine)
    ms, _, _ = time_graph(torch, dist, world, device, launch, steps, 3, d)
    us = ms * 1e3 / steps
    gbs = windows * (8 * n + 16) / (us * 1e-6) / 1e9
    rec = {"label": label, "sf": sf, "windows_per_step_per_gpu": windows, "steps": steps,
           "fine": None if fine is None else {"index0": fine[0], "step": fine[1]},
           "ms_per_step": us * 1e-3, "value": world * windows / us, "unit": "Msymbols/s",
           "roofline": {"achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak},
           "kernel": d.last_kernel, "indices_bit_exact": ok}
    d.close()
    del bufs, out
    torch.cuda.empty_cache()
    return rec


def cfg4_entry(torch, dist, m, world, rank, device, local, steps, peak):
    """BASELINE configs[3]: 256 concurrent channels x SF10, 1 024 windows per channel, the channels sharded
    over the ranks (256 / N whole channels per GPU, each channel an independent device buffer), ONE
    lorab200_demod_multi_dev launch per GPU per step; NCCL only carried the tables."""
    sf, wpc = CFG4_SF, CFG4_WINDOWS
    n = 1 << sf
    nch = CFG4_CHANNELS // world
    d = m.Demod(sf, device=local, flags=m.FLAG_INPUT_STABLE)
    step_bytes = nch * wpc * n * 8
    nbuf = n_buffers(step_bytes)
    segs = []
    for b in range(nbuf):
        # channel c of rank r uses seed c_global, a different stream per rotation buffer
        segs.append([synth_on_device(sf, wpc, seed=(rank * nch + c) + 100003 * b, device=device)[0]
                     for c in range(nch)])
    out = torch.empty((nch * wpc, 4), dtype=torch.int32, device=device)
    d.demod_multi_dev(segs[0], wpc, out=out)
    torch.cuda.synchronize()
    got = m.results_to_numpy(out)
    ok = True
    for c in (0, nch // 2, nch - 1):
        _, _, okc = oracle_check(m, d, sf, segs[0][c], got[c * wpc:(c + 1) * wpc], n_check=32)
        ok = ok and okc

    def launch(i, s):
        d.demod_multi_dev(segs[i % nbuf], wpc, out=out, stream=s)
    ms, _, _ = time_graph(torch, dist, world, device, launch, steps, 3, d)
    us = ms * 1e3 / steps
    gbs = nch * wpc * (8 * n + 16) / (us * 1e-6) / 1e9
    rec = {"label": "cfg4", "workload": f"{CFG4_CHANNELS} channels x SF={sf} x {wpc} windows, sharded "
                                        f"{nch} channels per GPU over {world} GPU(s), one multi-channel launch per GPU per step",
           "sf": sf, "channels_total": CFG4_CHANNELS, "channels_per_gpu": nch, "windows_per_channel": wpc,
           "steps": steps, "ms_per_step": us * 1e-3, "value": world * nch * wpc / us, "unit": "Msymbols/s",
           "scaling": "strong", "roofline": {"achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak},
           "kernel": d.last_kernel, "indices_bit_exact": bool(ok)}
    d.close()
    del segs, out
    torch.cuda.empty_cache()
    return rec


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import lora_sdr_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_local = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    affinity0 = os.sched_getaffinity(0)
    placement = {"bound": False, "why": "--no-numa-bind"} if args.no_numa_bind else \
        bind_to_gpu_numa(local, max(n_local, torch.cuda.device_count() if world > 1 else 1))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    m.build()
    sf, nw, K, W = args.sf, args.windows * args.channels, args.steps, args.warmup
    n = 1 << sf
    nch, wpc = args.channels, args.windows
    assert nch == 1 or args.fine_step is None, "--channels is the detector-only form"
    # the synthetic batches are resident and never rewritten: let consecutive launches overlap
    # their first loads with the previous launch's tail (LORAB200_FLAG_INPUT_STABLE)
    d = m.Demod(sf, device=local, flags=m.FLAG_SIMPLE_ONLY if args.simple else
                (m.FLAG_DEFAULT if args.no_input_stable else m.FLAG_INPUT_STABLE))

    # ---- tables: rank 0's copy is broadcast over NCCL (the only collective on this path) -------
    if world > 1:
        for which in (m.TABLE_UP, m.TABLE_DOWN, m.TABLE_FINE, m.TABLE_TWIDDLE):
            ptr, cnt = d.table_dev_ptr(which)
            tview = _as_tensor(ptr, cnt * 2, device)
            dist.broadcast(tview, src=0)
        torch.cuda.synchronize()
        d.tables_sync_from_dev()

    # ---- synthetic IQ resident in HBM; enough distinct buffers to defeat the L2 -----------------
    step_bytes = nw * n * 8
    nbuf = n_buffers(step_bytes, args.l2_resident)
    bufs, syms = [], []
    for b in range(nbuf):
        iq, s = synth_on_device(sf, nw, seed=0x10AA + 7919 * b + 104729 * rank, device=device)
        bufs.append(iq)
        syms.append(s)
    fine = None if args.fine_step is None else (1234, float(args.fine_step))
    # per-channel buffers (independent allocations) for the multi-channel form
    segs = [[bufs[b][c * wpc:(c + 1) * wpc].clone() for c in range(nch)] for b in range(nbuf)] if nch > 1 else None
    out = torch.empty((nw, 4), dtype=torch.int32, device=device)

    def launch(i, cuda_stream):
        if nch > 1:
            d.demod_multi_dev(segs[i % nbuf], wpc, out=out, stream=cuda_stream)
        else:
            d.demod_dev(bufs[i % nbuf], nw, out=out, stream=cuda_stream, fine=fine)
    torch.cuda.synchronize()

    # ---- correctness spot check on this rank (sampled windows vs the CPU oracle) -------------------
    parity = spot_check(m, d, sf, bufs[0], syms[0], out)

    # ---- device-resident timing: K steps in one CUDA graph ----------------------------------------------
    sampler = ClockSampler(local, period_s=0.005)
    ms_max, launches_per_graph, sustained = time_graph(torch, dist, world, device, launch, K, W, d, sampler,
                                                       sustain_s=0.0 if args.no_sweep else 0.6)
    value = world * nw * K / (ms_max * 1e-3) / 1e6
    clocks = sampler.summary()
    if sustained is not None:
        sustained["value"] = world * nw / (sustained["ms_per_step"] * 1e-3) / 1e6
        sustained["unit"] = "Msymbols/s"
        sustained["roofline_frac"] = nw * (8 * n + 16) / (sustained["ms_per_step"] * 1e-3) / 1e9 / peaks()[0]
        sustained["note"] = ("same graph replayed back to back; after ~70 ms the part reaches its software power "
                             "cap (clocks.reasons) and the SM clock drops: the K-step `value` above is the burst figure")
    headline_kernel = d.last_kernel

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------------
    Ke = args.e2e_steps or min(K, 100)
    host_bufs, pinned = [], []
    for b in range(min(2, nbuf)):
        # page-locked buffers from the library's own allocator (what a block's buffer manager uses)
        pa = m.PinnedArray((nw, n))
        pa.array[...] = bufs[b].cpu().numpy()
        pinned.append(pa)
        host_bufs.append(pa.array)
    host_out = np.empty(nw, m.RESULT_DTYPE)

    def host_step(i):
        hb = host_bufs[i % len(host_bufs)]
        if nch > 1:   # the host-buffer entry point is per channel (one block per channel)
            for c in range(nch):
                d.demod_host(hb[c * wpc:(c + 1) * wpc], wpc, out=host_out[c * wpc:(c + 1) * wpc])
        else:
            d.demod_host(hb, nw, out=host_out)

    for i in range(3):
        host_step(i)
    # The host is shared (other tenants' jobs take cores and memory bandwidth for milliseconds at a time, and a
    # Ke-step region is only ~30-130 ms): the Ke consecutive steps are timed three times, each between a
    # barrier + synchronize, max over ranks, and the best repetition is reported -- the same treatment the
    # reference arm gets (best of its K steps); every repetition is listed in e2e.repetitions.
    e2e_reps = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(Ke):
            host_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_reps.append(world * nw * Ke / float(te.item()) / 1e6)
    e2e_value = max(e2e_reps)
    # the records of the LAST timed e2e step against the CPU oracle (a sample of its windows):
    # the host-buffer entry point finalises with the host libm, so whole records must match
    last = (Ke - 1) % len(host_bufs)
    pick, exp, idx_ok = oracle_check(m, d, sf, bufs[last], host_out)
    e2e_check = {"windows_checked_vs_oracle": int(len(pick)), "indices_bit_exact": idx_ok,
                 "records_bit_exact": bool(host_out[pick].tobytes() == exp.tobytes()),
                 "decoded_equals_sent": float(np.mean(host_out["index"] == syms[last].cpu().numpy()))}

    # the host->device copy of a step's input ALONE (same page-locked buffer, no kernel, no D2H), every rank
    # copying at the same time: the ceiling the e2e leg sits under (one PCIe gen5 x16 link per GPU, and at
    # N = 8 the host's memory system feeding eight of them)
    src_t = torch.from_numpy(host_bufs[0].reshape(-1).view(np.float32))
    dst_t = torch.empty(src_t.shape, dtype=torch.float32, device=device)
    for _ in range(2):
        dst_t.copy_(src_t, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n_copies = 10
    t0 = time.perf_counter()
    for _ in range(n_copies):
        dst_t.copy_(src_t, non_blocking=True)
    torch.cuda.synchronize()
    tc = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    h2d_only_gbs = step_bytes * n_copies / float(tc.item()) / 1e9
    del dst_t, src_t

    # the same call with ordinary (pageable) host memory, as a framework's default buffers are:
    # the library stages it through its own pinned ring (filled by a few helper threads)
    pageable = np.array(host_bufs[0])
    keep = host_bufs
    host_bufs = [pageable]
    Kp = max(2, min(Ke, 10))
    host_step(0)
    pageable_reps = []
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(Kp):
            host_step(i)
        torch.cuda.synchronize()
        tp = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        pageable_reps.append(world * nw * Kp / float(tp.item()) / 1e6)
    e2e_pageable = max(pageable_reps)
    host_bufs = keep

    # ---- the other configurations the metric names (every rank takes part: max over ranks) ------------------
    peak, peak_src = peaks()
    sweep, cfg4 = [], None
    if not args.no_sweep and not args.simple and nch == 1 and fine is None and not args.l2_resident:
        host_np = np.array(host_bufs[0][: min(nw, 1 << 17)])      # keep a sample for the CPU baseline
        del pinned, host_bufs, keep, pageable, bufs, segs, out
        torch.cuda.empty_cache()
        common = (torch, dist, m, world, rank, device, local)
        # BASELINE configs[2]: "SF sweep 7-12, 16 384 windows each" (the headline above is configs[1], SF7 x 65 536)
        for s_sf, s_w, s_steps in ((7, 16384, 200), (8, 16384, 200), (9, 16384, 100), (10, 16384, 100), (11, 16384, 60), (12, 16384, 40)):
            if s_sf != sf or s_w != nw:
                sweep.append(sweep_entry(*common, s_sf, s_w, s_steps, None, peak, f"sf{s_sf}"))
        # the fine-tune multiply of the reference's steady-state data path (LoRaDemod.cpp:157-166):
        # constant index (_finefreqError == 0) and an index walking 1.28 entries per sample
        for s_sf, s_w, s_steps in ((7, 65536, 200), (10, 16384, 60), (12, 16384, 30)):
            sweep.append(sweep_entry(*common, s_sf, s_w, s_steps, (1234, 0.0), peak, f"sf{s_sf}_fine_const"))
            sweep.append(sweep_entry(*common, s_sf, s_w, s_steps, (1234, 1.28), peak, f"sf{s_sf}_fine_walk_1.28"))
        if CFG4_CHANNELS % world == 0:
            cfg4 = cfg4_entry(*common, 60, peak)
    else:
        host_np = np.array(host_bufs[0][: min(nw, 1 << 17)])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel -----------------------------------------------------------------
    alg_bytes = nw * (8 * n + 16)
    kernel_ms = ms_max / K
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": load_traffic(sf), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes, "kernel": headline_kernel}

    cpu = None
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, affinity0)   # the CPU baseline gets every core the process was given
        cpu = cpu_baseline(sf, host_np, len(affinity0))

    line = {
        "metric": METRIC, "value": value, "unit": "Msymbols/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(args),
        "parity_spot_check": parity,
        "clocks": clocks,
        "sustained": sustained,
        "e2e": {"value": e2e_value, "unit": "Msymbols/s", "h2d_bytes_per_step": step_bytes,
                "d2h_bytes_per_step": nw * 32, "steps": Ke, "repetitions": [round(v, 3) for v in e2e_reps],
                "check": e2e_check,
                "symbols_ok": bool(e2e_check["indices_bit_exact"]),
                "pageable_input_value": e2e_pageable,
                "h2d_copy_alone_gbs_per_gpu": h2d_only_gbs,
                "h2d_gbs_in_e2e_per_gpu": e2e_value / world * 1e6 * (8 * n) / 1e9,
                "api": "lorab200_demod_host (host buffers from lorab200_alloc_pinned; H2D + kernel + D2H + host finalise per step)"},
        "gpu_launches": int(launches_per_graph),
        "roofline": roof,
        "sweep": sweep,
        "cfg4": cfg4,
        "cpu_baseline": cpu,
        "host_placement": placement,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _as_tensor(ptr, nfloats, device):
    """Wrap a raw device pointer (a table inside the library) as a float32 torch tensor."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=device)


def spot_check(m, d, sf, iq, syms, out):
    """Demod one buffer on the GPU and compare a sample of windows against the CPU oracle (bit-exact indices)."""
    import torch

    nw = iq.shape[0]
    d.demod_dev(iq, nw, out=out)
    torch.cuda.synchronize()
    got = m.results_to_numpy(out)
    pick, _, idx_ok = oracle_check(m, d, sf, iq, got)
    sym_rate = float(np.mean(got["index"] == syms.cpu().numpy()))
    assert idx_ok, "GPU indices differ from the CPU oracle"
    return {"windows_checked_vs_oracle": int(len(pick)), "indices_bit_exact": idx_ok,
            "decoded_equals_sent": sym_rate}


def load_traffic(sf):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), if present."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(f"sf{sf}")
        except Exception:
            return None
    return None


if __name__ == "__main__":
    main()
