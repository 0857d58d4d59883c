#!/usr/bin/env python
"""bench.py — STFT frames/s of the phase-vocoder hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 4096 mono channels PER GPU, frame 1024 / hop 256,
pitchFactor 0.8.  One "step" == one batch of synthetic input == --calls-per-step (64) consecutive
hops of all 4096 channels == 64 process() calls == 64 launches of the fused kernel == 262144 STFT
frames (window -> FFT -> peak shift -> IFFT -> overlap-add, 12*N algorithmic bytes each, SURVEY.md
section 8d).  (Round 1 timed one call per step: at the driver's --steps 20 the timed region was then
0.35 ms and a fifth of it was the drain of the last launch, which nothing overlaps.)  State of one batch (32 MiB) would fit in the 126 MB L2, so the
timed loop rotates over `rotate` independent 4096-channel processor instances whose
combined state exceeds L2 (timing rule: inputs larger than L2).

Prints ONE JSON line (rank 0).  `value` = frames/s over all GPUs with inputs resident in
HBM; `e2e` = the same through the host-buffer C-ABI call (pvb_process_many, pinned host
memory, H2D + D2H inside the timed region); `roofline` = algorithmic GB/s of the fused
kernel against the measured HBM copy peak; `cpu_baseline` = the CPU oracle (a C port of
the reference JS; no JS engine exists in this image) on the box's host cores.

--impl reference times that CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "STFT frames/sec (1024-pt, hop 256) at 4096 ch; achieved HBM GB/s vs peak"


def workload_name(channels, frame, hop, pitch):
    tag = {(1024, 256, 4096, 0.8): " (BASELINE configs[1])", (2048, 512, 2048, 1.5): " (BASELINE configs[2]: 1024 stereo streams)",
           (1024, 256, 32768, 1.25): " (BASELINE configs[3], all channels on one GPU)",
           (2048, 128, 2048, 1.2): " (the reference's own frame / hop)",
           (1024, 256, 4096, 0.6): " (low end of the demo's pitch slider: DEEP instances)",
           (1024, 256, 4096, 0.5): " (bottom of the demo's pitch slider: DEEP instances)"}.get((frame, hop, channels, round(float(pitch), 3)), "")
    if not tag and channels == 8192 and hop * 4 == frame:
        tag = " (BASELINE configs[4] sweep)"
    return f"{channels} mono channels per GPU, frame {frame} / hop {hop}, pitchFactor {pitch}{tag}"
UNIT = "frames/s"
FRAME, HOP, CHANNELS, PITCH = 1024, 256, 4096, 0.8
L2_BYTES = 126e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--calls-per-step", type=int, default=64,
                    help="process() calls per step: one step = one batch of this many consecutive hops "
                         "of every channel (64 x 4096 frames at the default workload)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frame", type=int, default=FRAME)
    ap.add_argument("--hop", type=int, default=0, help="default: frame/4")
    ap.add_argument("--channels", type=int, default=CHANNELS, help="channels per GPU per step")
    ap.add_argument("--pitch", type=float, default=PITCH)
    ap.add_argument("--rotate", type=int, default=0,
                    help="processor instances rotated through (0: enough to exceed 2x L2)")
    ap.add_argument("--peak-guard", type=int, default=0, choices=[0, 1, 2, 3],
                    help="PVB_OPT_PEAK_GUARD of the timed handles: 0 default policy, 1 off, 2 always, 3 strict")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched (64 calls per launch) run")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short device-resident runs of the other BASELINE configs (N=1, default workload only)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-root-scatter", action="store_true",
                    help="N>1: skip the single-root mode (NCCL scatter of input slabs, gather of outputs)")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------
# clocks: NVML sampler thread running across the timed region
# ------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        self.samples, self.ok, self._stop = [], False, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:       # pragma: no cover
            self.err = repr(e)
        self.t0 = self.t1 = None
        self.th = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self.th:
            self.th.join(timeout=1.0)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvml unavailable"}
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= self.t1]
        note = None
        if not inside:            # timed region shorter than the sampling period
            inside = self.samples[-3:]
            note = "timed region shorter than sampler period; nearest samples used"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "note": "no samples"}
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [n for b, n in self.REASONS.items() if bits & b and n != "gpu_idle"]
        out = {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons,
               "samples": len(inside)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------
# CPU baseline: the oracle (C port of the reference JS) on the host cores
# ------------------------------------------------------------------------------------
def cpu_port_rate(frame, hop, pitch, channels, calls, threads):
    """frames/s of the CPU oracle: `channels` split over `threads` processor instances,
    `calls` process() calls each.  Returns (frames_per_s, seconds)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle_lib
    from phaze_b200 import signals

    oracle_lib.build()
    per = [channels // threads + (1 if i < channels % threads else 0) for i in range(threads)]
    per = [c for c in per if c > 0]
    procs = [oracle_lib.OracleProcessor(frame, hop, c) for c in per]
    nblk = 8
    blocks = [signals.channels(0, max(per), nblk * hop).reshape(max(per), nblk, hop)
              .transpose(1, 0, 2).copy()]
    blk = blocks[0]

    def work(i):
        p, c = procs[i], per[i]
        for k in range(calls):
            p.process_packed(blk[k % nblk, :c], pitch)     # ctypes releases the GIL

    # prime history so every frame is full
    with ThreadPoolExecutor(len(per)) as ex:
        list(ex.map(lambda i: [procs[i].process_packed(blk[k % nblk, :per[i]], pitch)
                               for k in range(frame // hop)], range(len(per))))
        t0 = time.perf_counter()
        list(ex.map(work, range(len(per))))
        dt = time.perf_counter() - t0
    for p in procs:
        p.close()
    return channels * calls / dt, dt


def cpu_baseline(frame, hop, pitch, seconds):
    import numpy as np
    threads = os.cpu_count() or 1
    channels = 16 * threads
    rate, _ = cpu_port_rate(frame, hop, np.float32(pitch), channels, 8, threads)     # calibration
    calls = max(8, int(rate * seconds / channels))
    rate, dt = cpu_port_rate(frame, hop, np.float32(pitch), channels, calls, threads)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{channels} channels x {calls} process() calls ({channels * calls} frames, "
                      f"{dt:.1f} s) of the {frame}/{hop} pf={pitch} workload; C port of the reference "
                      f"JS (oracle/phaze_oracle.c, -O2), one processor instance per thread"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    No JS engine exists in this image, so it is the oracle port (kind: port)."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frame, hop = args.frame, (args.hop or args.frame // 4)
    threads = os.cpu_count() or 1
    pitch = np.float32(args.pitch)
    # one step == calls_per_step process() calls over a bounded sample of the per-GPU batch
    cps = max(1, args.calls_per_step)
    probe, _ = cpu_port_rate(frame, hop, pitch, 8 * threads, 8, threads)
    budget_s = 120.0
    total = (args.steps + args.warmup) * cps
    chans = int(min(args.channels, max(threads, probe * budget_s / max(total, 1))))
    chans = max(threads, chans - chans % threads)
    if args.warmup:
        cpu_port_rate(frame, hop, pitch, chans, args.warmup * cps, threads)
    rate, dt = cpu_port_rate(frame, hop, pitch, chans, args.steps * cps, threads)
    sample = (f"each step = {cps} process() calls over {chans} of the {args.channels} channels "
              f"({frame}/{hop}, pf={args.pitch}); {threads} host threads, one processor instance each")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.channels, frame, hop, args.pitch),
                   "frame": frame, "hop": hop, "pitch_factor": args.pitch,
                   "channels_per_gpu": args.channels, "calls_per_step": cps, "channels_per_step": chans},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C port of the reference JS (oracle/phaze_oracle.c); Node/V8 is not present in this image",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    import phaze_b200
    from phaze_b200 import signals

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; phaze_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    # run this rank's host thread (and so the first touch of its pinned buffers) on the CPUs next to its
    # GPU: the end-to-end number is PCIe-bound (no effect on the single-node VMs of this pool, where the
    # e2e value still moved between 2.9e7 and 4.15e7 frames/s from box to box with the same code)
    orig_affinity = os.sched_getaffinity(0)
    numa = {"bound": False}
    try:
        if os.environ.get("PVB_BENCH_NO_NUMA") == "1":
            raise RuntimeError("disabled")
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = {"bound": True, "cpus": len(os.sched_getaffinity(0)), "of": len(orig_affinity)}
    except Exception as e:       # containers without the permission: keep the inherited affinity
        numa = {"bound": False, "why": type(e).__name__}
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    frame, hop = args.frame, (args.hop or args.frame // 4)
    C, K, W = args.channels, args.steps, args.warmup
    CPS = max(1, args.calls_per_step)
    R = frame // hop
    pitch = np.float32(args.pitch)
    state_bytes = 2 * C * frame * 4 + 2 * C * hop * 4
    rotate = args.rotate or max(2, int(np.ceil(2.2 * L2_BYTES / state_bytes)))
    first_channel = rank * C          # weak scaling: every rank owns its own channel range

    # synthetic input: distinct blocks per channel covering at least two frames (the blocks are reused
    # cyclically: a period shorter than the frame would be a line spectrum with most bins at the round-off
    # floor -- ill-conditioned input, which the kernels answer with their float64 peak decisions), resident
    # in HBM before timing
    nblk = max(8, 2 * R)
    host = signals.channels(first_channel, C, nblk * hop)
    blocks_np = np.ascontiguousarray(host.reshape(C, nblk, hop).transpose(1, 0, 2))
    blocks = torch.from_numpy(blocks_np).cuda()
    outs = [torch.empty((C, hop), dtype=torch.float32, device="cuda") for _ in range(rotate)]
    # inputs_ready: the input blocks are resident in HBM before anything is timed (the library's default
    # is strict stream order: the first launch of every submission waits for the whole stream, because
    # its input could come from a kernel the library knows nothing about)
    procs = [phaze_b200.BatchedPhaseVocoder(C, frame, hop, device=local, inputs_ready=1, peak_guard=args.peak_guard)
             for _ in range(rotate)]
    # a real (non-NULL) stream: NULL would select the handle's own stream and the events
    # below would not bracket the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    def call(i):
        procs[i % rotate].process_device(blocks[i % nblk].data_ptr(), outs[i % rotate].data_ptr(),
                                         pitch, sptr)

    def step(i):
        # one step: CPS consecutive process() calls, round-robin over the resident processor instances
        for j in range(i * CPS, (i + 1) * CPS):
            call(j)

    # prime: fill every instance's history so all frames carry signal
    for i in range(rotate * R):
        call(i)
    torch.cuda.synchronize()

    # The clock sampler runs on rank 0 only (one NVML thread per node instead of one per GPU)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    def timed_run(steps, warm):
        """barrier -> gate -> `warm` untimed launches -> ev0 -> `steps` launches -> ev1, all queued
        back to back on one stream.  The gate (a ~1.5 ms device-side sleep in front of the warm-up)
        lets the host enqueue the whole sequence before the GPU starts on it, so even a 20-step run
        measures the launches at their steady-state spacing instead of the host's enqueue rate or a
        cold start after a synchronize(); nothing but our own launches sits between the two events."""
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(3e6))
        for i in range(warm):
            step(i)
        e0.record(stream)
        for i in range(steps):
            step(warm + i)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    timed_run(max(W, 3), 0)            # untimed: first-use costs (module load, allocator, clocks ramp)
    launches0 = sum(p.kernel_launches for p in procs)
    guard0 = sum(p.peak_guard_count for p in procs)
    if sampler:
        sampler.mark_begin()
    ms = timed_run(K, W)
    if sampler:
        sampler.mark_end()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    launches = sum(p.kernel_launches for p in procs) - launches0 - W * CPS
    guard_frames = sum(p.peak_guard_count for p in procs) - guard0
    if sampler:
        sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    frames_total = world * K * CPS * C
    value = frames_total / (ms_max * 1e-3)

    # the same K steps with the launch chaining switched off (PVB_OPT_LAUNCH_MODE): 1 = programmatic
    # dependent launch with a whole-grid wait, 2 = plain stream-ordered launches.  The difference to
    # `value` is what the overlap of consecutive launches buys (evidence for the roofline figure: an
    # ncu launch list times every launch alone and cold, like mode 2).
    chain = {}
    for mode, name in ((1, "pdl_grid_wait"), (2, "plain_launches")):
        for p in procs:
            p.set_option("launch_mode", mode)
        ms_m = timed_run(K, W)
        chain[name] = {"value": K * CPS * C / (ms_m * 1e-3), "avg_launch_us": 1e3 * ms_m / (K * CPS)}
    for p in procs:
        p.set_option("launch_mode", 0)

    # batched mode (pvb_process_many_device): Kb consecutive calls per launch, every
    # channel pair loops over them, its state goes through L1 / L2 instead of HBM.  Compulsory DRAM traffic
    # drops to 8*hop bytes per frame (+ 12*N / Kb), so this is NOT scored against the 12*N roofline: it is
    # reported as frames/s and against the non-tensor FP32 ceiling (BASELINE.md contract).
    batched = None
    if not args.no_batched:
        Kb = 64
        bin_ = blocks.repeat((Kb + nblk - 1) // nblk, 1, 1)[:Kb].contiguous()        # [Kb][C][hop]
        bouts = [torch.empty((Kb, C, hop), dtype=torch.float32, device="cuda") for _ in range(2)]
        nb = max(2, min(rotate, K * CPS // Kb))

        for p in procs:
            p.set_option("many_mode", 1)

        def bstep(i):
            procs[i % rotate].process_device(bin_.data_ptr(), bouts[i % 2].data_ptr(), pitch, sptr, num_calls=Kb)

        for i in range(rotate):
            bstep(i)
        torch.cuda.synchronize()
        eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eb0.record(stream)
        for i in range(nb):
            bstep(i)
        eb1.record(stream)
        torch.cuda.synchronize()
        bms = eb0.elapsed_time(eb1)
        bval = nb * Kb * C / (bms * 1e-3)
        flop_per_frame = 6.0e4 * frame / 1024          # ~60 k float32 flop per 1024-point frame
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12     # non-tensor FP32, TFLOP/s at the maximum SM clock
        batched = {"value": bval, "unit": UNIT, "calls_per_launch": Kb, "launches": nb,
                   "us_per_call": 1e3 * bms / (nb * Kb),
                   "vs_one_launch_per_call": bval / (K * CPS * C / (ms * 1e-3)),
                   "compulsory_dram_bytes_per_frame": 8 * hop + 12.0 * frame / Kb,
                   "fp32": {"flop_per_frame_estimate": flop_per_frame, "achieved_tflops": bval * flop_per_frame / 1e12,
                            "peak_tflops": fp32_peak, "frac": bval * flop_per_frame / 1e12 / fp32_peak},
                   "api": f"pvb_process_many_device(handle, in_dev, out_dev, {Kb}, pitch, stream) with "
                          "PVB_OPT_MANY_MODE = 1: one launch, bit-identical to the same calls one by one"}
        del bin_, bouts
        for p in procs:
            p.set_option("many_mode", 0)

    # context only: (a) one instance, state stays in L2; (b) every instance on its own stream
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    KC = K * CPS                       # calls of the context runs
    for i in range(50):
        procs[0].process_device(blocks[i % nblk].data_ptr(), outs[0].data_ptr(), pitch, sptr)
    ev2.record(stream)
    for i in range(KC):
        procs[0].process_device(blocks[i % nblk].data_ptr(), outs[0].data_ptr(), pitch, sptr)
    ev3.record(stream)
    torch.cuda.synchronize()
    l2_value = KC * C / (ev2.elapsed_time(ev3) * 1e-3)

    side = [torch.cuda.Stream() for _ in range(rotate)]
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev4.record(stream)
    for st in side:
        st.wait_event(ev4)
    for i in range(KC):
        procs[i % rotate].process_device(blocks[i % nblk].data_ptr(), outs[i % rotate].data_ptr(),
                                         pitch, side[i % rotate].cuda_stream)
    for st in side:
        e = torch.cuda.Event()
        e.record(st)
        stream.wait_event(e)
    ev5.record(stream)
    torch.cuda.synchronize()
    streams_value = KC * C / (ev4.elapsed_time(ev5) * 1e-3)

    # e2e: host buffers through the C ABI (pinned), H2D + D2H in the timed region
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(procs, blocks_np, C, hop, pitch, K, CPS, dist)

    root_scatter = None
    if dist and not args.no_root_scatter:
        root_scatter = run_root_scatter(C, frame, hop, pitch, world, rank, local, min(KC, 300), host)

    peak, peak_src = measured_peak()
    traffic, traffic_src = (ncu_traffic_bytes() if (frame, hop, C, round(float(pitch), 3)) == (FRAME, HOP, CHANNELS, PITCH)
                            else (None, None))
    assert launches == K * CPS, (launches, K, CPS)
    kernel_ms = ms / max(launches, 1)                 # this rank's average launch duration (start to start)
    achieved = 12.0 * frame * C / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                # dram__bytes_read + dram__bytes_write per launch, parsed from the newest committed
                # `ncu --set full` summary of this kernel (default workload only)
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel": procs[0].kernel_name(pitch),
                "algorithmic_bytes_per_launch": 12 * frame * C,
                "avg_launch_us": kernel_ms * 1e3}

    # the other BASELINE configs (parity-test cases, not bench lines): short device-resident runs on
    # the same terms as `value`, reported as context beside the headline
    others = None
    is_default = (frame, hop, C, round(float(args.pitch), 3)) == (FRAME, HOP, CHANNELS, PITCH)
    if world == 1 and is_default and not args.no_other_configs:
        for p in procs:
            p.close()
        procs = []
        del outs, blocks
        torch.cuda.empty_cache()
        peak_o, _ = measured_peak()
        others = [quick_config(local, f, h, c, pf, peak_o) for (f, h, c, pf) in OTHER_CONFIGS]

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, orig_affinity)      # the CPU baseline uses every host core
            cpu = cpu_baseline(frame, hop, args.pitch, args.cpu_seconds)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(C, frame, hop, args.pitch),
                       "frame": frame, "hop": hop, "pitch_factor": args.pitch,
                       "channels_per_gpu": C, "calls_per_step": CPS, "frames_per_step": C * CPS * world,
                       "step": f"one batch of {CPS} consecutive hops of every channel = {CPS} process() calls = "
                               f"{CPS} kernel launches (round-robin over the resident processor instances)",
                       "l2_policy": f"inputs larger than L2: {rotate} processor instances "
                                    f"({rotate * state_bytes / 2**20:.0f} MiB of state+io) rotated per call",
                       "parallelism": f"channel-sharded x{world}, no data-path collective",
                       "launch": "one stream; consecutive launches chained by programmatic dependent launch, "
                                 "each channel pair waits for its own previous call (completion flags)",
                       "timed_region": "barrier, device-side gate, W warm-up steps, event, K steps, event: "
                                       "queued back to back, no synchronize between warm-up and timed steps"},
            "roofline": roofline,
            "e2e": e2e,
            "host_thread_affinity": numa,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "launch_chaining": chain,
            "batched": batched,
            "peak_guard": {"option": args.peak_guard, "frames_redecided_in_float64": int(guard_frames),
                           "of_frames": int((K + W) * CPS * C),
                           "note": "channel frames whose peak set was re-decided with the fft.js-order float64 "
                                   "transform during warm-up + timed launches (this rank)"},
            "l2_resident_value": l2_value,
            "concurrent_streams_value": streams_value,
            "cpu_baseline": cpu,
        }
        if root_scatter:
            line["root_scatter"] = root_scatter
        if others:
            line["other_configs"] = others
        print(json.dumps(line))
    for p in procs:
        p.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(procs, blocks_np, C, hop, pitch, K, CPS, dist):
    """Same metric through the host-buffer entry points: every step copies its input block from
    pinned host memory to the device and its output block back (4 MiB each way at the default
    workload).  `value` uses pvb_process_many (128 consecutive process() calls per submission,
    copies and kernels pipelined on three streams, bit-identical to 128 single calls);
    `single_call_value` uses the synchronous one-call-at-a-time pvb_process."""
    import ctypes as Ct

    import numpy as np
    import torch

    lib = phaze_b200_lib()
    nblk = blocks_np.shape[0]
    nbytes = C * hop * 4
    batch = CPS                       # one submission == one step: CPS process() calls
    hin = lib.pvb_alloc_host(nbytes * batch)
    hout = lib.pvb_alloc_host(nbytes * batch)
    for k in range(batch):
        Ct.memmove(hin + k * nbytes, blocks_np[k % nblk].ctypes.data, nbytes)

    def timed(fn, steps, per_call):
        # warm-up: one submission per rotated handle (the first one allocates the handle's device
        # staging buffers, streams and events: set-up, not steady state)
        for i in range(len(procs)):
            fn(i)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps // per_call):
            fn(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        world = 1
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            world = dist.get_world_size()
        return world * (steps // per_call) * per_call * C / float(t.item())

    def many(i):
        rc = lib.pvb_process_many(procs[i % len(procs)]._h, hin, hout, batch, pitch)
        assert rc == 0, rc

    def single(i):
        rc = lib.pvb_process(procs[i % len(procs)]._h, hin + (i % batch) * nbytes, hout, pitch)
        assert rc == 0, rc

    steps = int(min(K, 64)) * batch                   # process() calls: K steps (at most 64) of `batch` calls
    value = timed(many, steps, batch)
    single_value = timed(single, int(min(K * CPS, 400)), 1)
    check = float(np.ctypeslib.as_array(Ct.cast(hout, Ct.POINTER(Ct.c_float)), (C * hop,)).std())
    # what bounds it: the host link.  Probe: the same two pinned buffers copied both ways at once on
    # two streams, nothing else running (GB/s in each direction)
    dev_a = torch.empty(nbytes * batch // 4, dtype=torch.float32, device="cuda")
    dev_b = torch.empty(nbytes * batch // 4, dtype=torch.float32, device="cuda")
    rt = Ct.CDLL("libcudart.so.12")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both_ways():
        rt.cudaMemcpyAsync(Ct.c_void_p(dev_a.data_ptr()), Ct.c_void_p(hin), Ct.c_size_t(nbytes * batch), 1, Ct.c_void_p(s1.cuda_stream))
        rt.cudaMemcpyAsync(Ct.c_void_p(hout), Ct.c_void_p(dev_b.data_ptr()), Ct.c_size_t(nbytes * batch), 2, Ct.c_void_p(s2.cuda_stream))
    probe = None
    try:
        both_ways()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            both_ways()
        torch.cuda.synchronize()
        probe = 4 * nbytes * batch / (time.perf_counter() - t0) / 1e9
    except Exception:       # pragma: no cover
        pass
    world = dist.get_world_size() if dist else 1
    lib.pvb_free_host(hin)
    lib.pvb_free_host(hout)
    return {"value": value, "unit": UNIT, "h2d_bytes_per_step": nbytes * batch, "d2h_bytes_per_step": nbytes * batch,
            "steps": steps // batch, "calls_per_submission": batch,
            "api": f"pvb_process_many(handle, in_host, out_host, {batch}, pitch): pinned host buffers, "
                   "H2D / kernel / D2H of consecutive calls overlapped, synchronous on return",
            "single_call_value": single_value, "out_std": check,
            "bound": "pcie (host link): every frame moves hop*4 bytes each way",
            "gbs_each_way_per_gpu": value / world * hop * 4 / 1e9,
            "pcie_probe_gbs_each_way_per_gpu": probe,
            "pcie_probe": "the same pinned buffers copied H2D and D2H concurrently, all ranks at once, no kernels"}


# BASELINE configs 3, 4 (one shard's worth per launch) and 5, plus the reference's own 2048 / 128
OTHER_CONFIGS = [(2048, 512, 2048, 1.5), (1024, 256, 32768, 1.25),
                 (256, 64, 8192, 1.2), (512, 128, 8192, 1.2), (1024, 256, 8192, 1.2), (2048, 512, 8192, 1.2),
                 (4096, 1024, 8192, 1.2), (2048, 128, 2048, 1.2),
                 # the low end of the demo's pitch slider (www/index.html:23): the ring-order kernel's DEEP instances
                 (1024, 256, 4096, 0.6), (1024, 256, 4096, 0.5)]


def quick_config(local, frame, hop, C, pf, peak, steps=300, warm=40, **options):
    """Device-resident throughput of one configuration, measured like `value`: instances rotated so
    that state + io exceed L2, one stream, CUDA events around `steps` launches."""
    import numpy as np
    import torch

    import phaze_b200
    from phaze_b200 import signals

    pitch = np.float32(pf)
    state_bytes = 2 * C * frame * 4 + 2 * C * hop * 4
    rotate = max(2, int(np.ceil(2.2 * L2_BYTES / state_bytes)))
    nblk = max(4, 2 * (frame // hop))          # the cyclic input must not have a period shorter than the frame
    host = signals.channels(0, C, nblk * hop)
    blocks = torch.from_numpy(np.ascontiguousarray(host.reshape(C, nblk, hop).transpose(1, 0, 2))).cuda()
    outs = [torch.empty((C, hop), dtype=torch.float32, device="cuda") for _ in range(rotate)]
    procs = [phaze_b200.BatchedPhaseVocoder(C, frame, hop, device=local, inputs_ready=1, **options) for _ in range(rotate)]
    stream = torch.cuda.Stream()        # (stream 0 would mean "the handle's own stream" to the library)
    sptr = stream.cuda_stream
    torch.cuda.synchronize()

    def step(i):
        procs[i % rotate].process_device(blocks[i % nblk].data_ptr(), outs[i % rotate].data_ptr(), pitch, sptr)

    for i in range(rotate * (frame // hop)):
        step(i)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        torch.cuda._sleep(int(3e6))    # gate: the whole sequence is queued before the GPU starts on it
    for i in range(warm):
        step(i)
    ev0.record(stream)
    for i in range(steps):
        step(warm + i)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    kernel = procs[0].kernel_name(pitch)
    for p in procs:
        p.close()
    achieved = 12.0 * frame * C / (ms * 1e-3) / 1e9
    return {"workload": workload_name(C, frame, hop, pf), "value": C / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms, "steps": steps, "roofline_frac": achieved / peak, "achieved_gbs": achieved,
            "kernel": kernel, "instances_rotated": rotate}


def run_root_scatter(C, frame, hop, pitch, world, rank, local, steps, host_block_src):
    """Single-root mode of the north-star: rank 0 holds the [world*C][hop] block in HBM, scatters the
    slabs over NVLink (grouped ncclSend/ncclRecv), every rank runs its shard, outputs are gathered
    back on rank 0.  Reports whole-job frames/s and the bytes that cross NVLink per second."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from phaze_b200.sharded import ShardedPhaseVocoder

    total = world * C
    sh = ShardedPhaseVocoder(total, frame, hop, device=torch.device("cuda", local))
    blk = None
    if rank == 0:
        blk = torch.from_numpy(np.tile(host_block_src[:, :hop], (world, 1))).cuda().contiguous()
    for _ in range(8):
        sh.process_from_root(blk, pitch)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        sh.process_from_root(blk, pitch)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item()) * 1e-3
    peer_bytes = 2 * (total - C) * hop * 4                     # slabs out + results back, per step
    res = {"value": steps * total / sec, "unit": UNIT, "steps": steps, "total_channels": total,
           "nvlink_bytes_per_step": peer_bytes, "nvlink_gbs_at_root": steps * peer_bytes / sec / 1e9,
           "nvlink_peer_copy_reference_gbs": 770.0,
           "note": "root egress + ingress; shard-resident number is the main `value`"}
    del sh
    # streamed root mode: audio buffers of Kc calls per message ([C][Kc*hop]), scatter of buffer i+1
    # and gather of buffer i-1 under the kernels of buffer i
    Kc, nbuf = int(os.environ.get("PVB_BENCH_STREAM_CALLS", "64")), 6
    sh = ShardedPhaseVocoder(total, frame, hop, device=torch.device("cuda", local))
    bufs = None
    if rank == 0:
        one = torch.from_numpy(np.tile(host_block_src[:, :hop], (world, Kc))).cuda().contiguous()
        bufs = [one.clone() for _ in range(nbuf)]
    sh.process_stream_from_root(bufs, pitch, Kc, 3)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    reps = 4            # (one pass of 6 buffers is a 6 ms window: two modes, 1.7e8 / 2.9e8 on 2 GPUs, from run to run)
    ev0.record()
    for _ in range(reps):
        sh.process_stream_from_root(bufs, pitch, Kc, nbuf)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item()) * 1e-3 / reps
    res["streamed"] = {"value": nbuf * Kc * total / sec, "unit": UNIT, "calls_per_message": Kc, "buffers": nbuf,
                       "passes_timed": reps,
                       "message_bytes_per_peer": C * Kc * hop * 4,
                       "nvlink_gbs_at_root": nbuf * Kc * peer_bytes / sec / 1e9,
                       "api": "ShardedPhaseVocoder.process_stream_from_root"}
    return res


def phaze_b200_lib():
    import phaze_b200
    return phaze_b200.load_library()


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel, parsed from the
    newest committed `ncu --set full` summary under profiles/ (r<NN>_ncu_ring_1024.txt, written by
    profiles/ncu_summary.py from a capture of this bench command).  (None, None) when there is none."""
    import glob
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_ring_1024.txt")), reverse=True):
        total, seen = 0.0, 0
        with open(path) as f:
            for ln in f:
                m = re.match(r"dram__bytes_(read|write)\.sum\s+(\w+)\s+([0-9.eE+-]+)", ln)
                if m and m.group(2) in unit:
                    total += float(m.group(3)) * unit[m.group(2)]
                    seen += 1
                if seen == 2:
                    return total, os.path.relpath(path, ROOT)
    return None, None


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
