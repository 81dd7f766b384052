#!/usr/bin/env python
"""Benchmark of the NWS forward hot path (BASELINE.json metric: audio samples/sec at 16 kHz).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--variant fastnewt|newt] [--batch-per-gpu B] [--seconds S]

A step = one NeuralWaveshaping.forward over one batch of synthetic control streams.  N=1 workload
(default): BASELINE.json configs[1] — FastNEWT LUT path, batch 64 x 4 s @ 16 kHz, inputs
torch.rand like scripts/time_forward_pass.py:27-40, random-init weights of newt.gin
(torch.manual_seed(0)); `--variant newt` gives configs[2].  N>1: the same batch per GPU (weak
scaling), one process per GPU under torchrun, no collective on the data path, one NCCL
all-reduce/all-gather of (seconds, samples) at the end.

The same line carries a `configs` object with the other BASELINE.json configurations, each with its own parity figure
against tests/golden: `c3` (configs[2]: full NEWT shapers, 64 x 4 s, own roofline), `c4` (configs[3]: B = 1 stateless
forward and stateful SynthStream.push at the buffer sizes of scripts/time_buffer_sizes.py:13, next to the CPU port at
the same sizes) and `c5` (configs[4]: a GLOBAL batch of 2048 utterances split over the ranks with shard_bounds —
256 per GPU at N = 8 — forward plus a timed NCCL gather of the audio).  `value` stays configs[1] so the series is comparable.

One JSON line on stdout (rank 0).  `value` is device-timed: the K steps are issued as a throughput job — two forwards in
flight on two streams / engines, as the public streaming.HostPipeline runs them — between two CUDA events on the launch
stream, inputs resident in HBM and rotating through more distinct batches than the L2 holds; `latency` is the same
forward one at a time with the L2 flushed before each (round 1's figure).  `e2e` is the same job through
the public API (HostPipeline) from pinned host tensors with the H2D/D2H copies inside the timed region;
`roofline` times the dominant kernel (nws_audio_fused_kernel) with the library's own stage events;
`cpu_baseline` / `--impl reference` time the oracle port (the reference's op sequence on torch CPU).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

SR, HOP = 16000, 128
ALGO_BYTES_PER_UTT_FRAME = 4 + 8 + 512   # f0 (1 fp32) + control (2 fp32) read, 128 fp32 samples written per frame


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="fastnewt", choices=["fastnewt", "newt"])
    ap.add_argument("--batch-per-gpu", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--cpu-batch", type=int, default=8, help="utterances per CPU-reference step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs.c3/c4/c5 legs (quick runs)")
    ap.add_argument("--c5-global-batch", type=int, default=2048)
    ap.add_argument("--inputs", default="rand", choices=["rand", "realistic"],
                    help="rand = the reference timing scripts' torch.rand f0/control (the contract's workload); "
                         "realistic = violin checkpoint + vibrato around 110-660 Hz (SURVEY.md 8(d)'s second input set)")
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions: NVML polled from a thread every ~2 ms (a 20-step run is
    only ~100 ms long: `nvidia-smi -lms 20` gave 7 samples), falling back to an nvidia-smi loop if NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread, self.stop_flag, self.source = index, [], None, None, False, None
        self.period_s = 0.002

    def _nvml_loop(self, pynvml, h):
        # NVML clocks-event-reason bits (nvml.h): HwSlowdown 0x8, HwThermalSlowdown 0x40, SwThermalSlowdown 0x20, SwPowerCap 0x4
        bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")]
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = get_reasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b, _ in bits])
            except Exception:
                pass
            time.sleep(self.period_s)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.source = "nvml, 2 ms period"
            self.thread = threading.Thread(target=self._nvml_loop, args=(pynvml, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 20"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(self.NAMES, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def build_weights():
    """gin-configured random-init weights, torch.manual_seed(0)."""
    import gin
    import torch
    from neural_waveshaping_synthesis.models.neural_waveshaping import NeuralWaveshaping
    gin.clear_config()
    gin.parse_config_file(os.path.join(REPO, "gin", "models", "newt.gin"))
    torch.manual_seed(0)
    return NeuralWaveshaping().eval()


def cpu_reference_throughput(model, variant: str, T: int, batch: int, steps: int, warmup: int, budget_s=None):
    """Times the oracle port (reference op sequence, torch CPU, all host threads).  Returns
    (samples_per_s, ms_per_step, threads)."""
    import torch
    from oracle import nws_oracle as oracle
    w = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    lut = oracle.build_lookup_table(w) if variant == "fastnewt" else None
    torch.manual_seed(1)
    f0 = torch.rand(batch, 1, T)
    control = torch.rand(batch, 2, T)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        u, noise = oracle.draw_rng(T)
        oracle.forward(w, f0, control, u, noise, lut=lut, faithful_loop=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and times and time.perf_counter() - t_start > budget_s:
            break   # bounded: the CPU arm must end within minutes whatever K is (steps actually timed are reported)
    mean = sum(times) / len(times)
    cpu_reference_throughput.steps_timed = len(times)
    return batch * T * HOP / mean, mean * 1e3, torch.get_num_threads()


def best_cpu_baseline(cpu_model, args, T):
    """The reference's CPU path on this box: torch's intra-op thread count matters a lot for these small
    ops (all cores is far from the best on a 100+ core host), so a few counts are tried and the best
    throughput is reported — the comparison is against the reference at its best."""
    import torch
    ncpu = os.cpu_count() or 1
    tried = []
    for threads in sorted({min(ncpu, t) for t in (8, 16, 32)}):   # all-cores is pathological (measured 100x slower at 128)
        torch.set_num_threads(threads)
        for batch, iters in ((1, 3), (args.cpu_batch, 2)):
            sps, _, _ = cpu_reference_throughput(cpu_model, args.variant, T, batch, iters, 1)
            tried.append((sps, threads, batch))
    sps, threads, batch = max(tried)
    return {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": "oracle port (reference op sequence incl. the faithful _lookup loop, torch CPU), %g s utterances; "
                      "best of threads x batch: %s -> %d threads, batch %d" %
                      (args.seconds, ", ".join("%dt/B%d: %.3g" % (t, b, v) for v, t, b in tried), threads, batch)}


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return
    import torch
    T = int(SR * args.seconds) // HOP
    model = build_weights()
    # pick the intra-op thread count at which the reference's CPU path runs fastest on this host
    best = None
    for threads in sorted({min(os.cpu_count() or 1, t) for t in (8, 16, 32)}):
        torch.set_num_threads(threads)
        probe, _, _ = cpu_reference_throughput(model, args.variant, T, args.cpu_batch, 1, 1)
        if best is None or probe > best[0]:
            best = (probe, threads)
    torch.set_num_threads(best[1])
    sps, ms, threads = cpu_reference_throughput(model, args.variant, T, args.cpu_batch, args.steps, args.warmup, budget_s=120.0)
    sample = ("oracle port of NeuralWaveshaping.forward (%s, faithful _lookup loop), %d of the %d utterances per step, "
              "%g s each, torch CPU %d threads" % (args.variant, args.cpu_batch, args.batch_per_gpu * args.gpus,
                                                  args.seconds, threads))
    line = {
        "impl": "reference", "metric": "audio samples/sec", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "steps_timed": cpu_reference_throughput.steps_timed, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf_per_utterance": (ms / 1e3) / (args.cpu_batch * args.seconds),
        "config": workload_config(args),
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def profile_path(variant):
    """The committed `ncu --set full` summary of the dominant kernel for this variant: newest round first."""
    tag = "lut" if variant == "fastnewt" else "mlp"
    for rnd in ("r2", "r1"):
        path = os.path.join(REPO, "profiles", "%s_ncu_audio_tc_%s.json" % (rnd, tag))
        if os.path.exists(path):
            return path
    return os.path.join(REPO, "profiles", "r1_ncu_audio_tc_%s.json" % tag)


def ncu_traffic(variant):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/, written by scripts/ncu_summary.py); None if absent."""
    path = profile_path(variant)
    try:
        d = json.load(open(path))[0]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        return rd["value"] * scale[rd["unit"]] + wr["value"] * scale[wr["unit"]]
    except Exception:
        return None


def issue_view(variant, kernel_ms=None, clocks=None, ffma_tflops=None):
    """What actually bounds the fused kernel (it is neither HBM- nor tensor-bound): warp-instruction issue.
    Instruction counts and pipe activity from the committed ncu capture of this workload (profiles/); the kernel
    time, the SM clock and the fp32 FMA rate are measured live, so `frac_of_issue_peak` = warp-instructions per
    second ÷ (SMs x 4 schedulers x SM clock) is this run's fraction of the issue roofline."""
    path = profile_path(variant)
    try:
        d = json.load(open(path))[0]
        g = lambda k: d[k]["value"]
        v = {"bound": "warp-instruction issue (fp32 SIMT epilogue + sine generation)",
             "counts_from": "committed ncu capture (not measured in this run): profiles/" + os.path.basename(path),
             "issue_slot_utilisation": g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
             "tensor_pipe_active": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
             "warp_instructions_per_launch": g("smsp__inst_executed.sum"), "source": os.path.basename(path)}
        n_sm = int(g("launch__grid_size"))   # persistent kernel: one CTA per SM
        if kernel_ms and clocks and clocks.get("sm_mhz"):
            rate = v["warp_instructions_per_launch"] / (kernel_ms * 1e-3)
            peak = n_sm * 4 * clocks["sm_mhz"] * 1e6
            v.update({"warp_instructions_per_s": rate, "issue_peak_warp_instructions_per_s": peak,
                      "frac_of_issue_peak": rate / peak})
        if ffma_tflops:
            v["fp32_ffma_tflops_measured"] = ffma_tflops
        return v
    except Exception:
        return {"bound": "warp-instruction issue", "source": None}


BUFFER_SIZES = [256, 512, 1024, 2048, 4096, 8192, 16384, 32768]   # scripts/time_buffer_sizes.py:13
STREAM_SIZES = [256, 512, 1024, 2048, 4096]                       # BASELINE.json configs[3]'s range


def golden_case(name, prefix=""):
    """A fixture of tests/golden (made by the real reference, oracle/gen_golden.py) as torch tensors; the noise vector is
    regenerated with the reference's draw order (torch CPU generator: rand(1,101,1), then rand(128T-1)) and the stored
    phase draw guards against RNG drift.  numpy + torch only: the GPU arm never imports oracle/."""
    import numpy as np
    import torch
    z = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
    c = {k[len(prefix):]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(prefix)}
    T = c["f0"].shape[-1]
    torch.manual_seed(int(c["rng_seed"]))
    u = torch.rand(1, 101, 1)
    c["noise"] = torch.rand(HOP * T - 1)
    if not torch.equal(u.reshape(-1), c["u_phase"]):
        raise RuntimeError("torch CPU RNG stream changed: golden noise cannot be regenerated")
    return c


def parity_max_abs(model, case, dev, out_key="out"):
    import torch
    with torch.no_grad():
        y = model(case["f0"].to(dev), case["control"].to(dev), phase_shift=case["u_phase"].to(dev), noise=case["noise"].to(dev))
    return float((y.cpu().double() - case[out_key].double()).abs().max())


LANES = 2


def lane_throughput_ms(model, dev, make_inputs, k, n_sets):
    """K steps as a throughput job: forwards issued alternately on LANES compute streams / engines (the public
    streaming.HostPipeline does the same with host buffers), inputs rotating through `n_sets` distinct device-resident
    batches (more bytes than the L2 holds, so no step finds its inputs cached; a flush between steps would serialise
    them).  The timed region is bracketed by events on the current stream: every lane waits for the first and the second
    waits for every lane.  Returns ms per step."""
    import torch
    sets = [make_inputs(i) for i in range(n_sets)]
    streams = [torch.cuda.Stream(dev) for _ in range(LANES)]
    outs = [None] * (2 * LANES)
    cur = torch.cuda.current_stream(dev)

    def issue(n, first):
        for i in range(n):
            f0, control = sets[(first + i) % n_sets]
            with torch.cuda.stream(streams[i % LANES]):
                outs[i % len(outs)] = model.forward_lane(i % LANES, f0, control, out=outs[i % len(outs)])

    with torch.no_grad():
        for st in streams:
            st.wait_stream(cur)
        issue(2 * LANES + 2, 0)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for st in streams:
            st.wait_event(a)
        issue(k, 7)
        for st in streams:
            cur.wait_stream(st)
        b.record()
        torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / k


def pctl(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


def config_c3(args, cpu_model, dev, timed_steps, flush, peaks_hbm, clocks_now):
    """BASELINE.json configs[2]: full NEWT sine-MLP shapers, batch 64 x 4 s, one GPU."""
    import copy
    import torch
    model = copy.deepcopy(cpu_model).to(dev)
    B, T = args.batch_per_gpu, int(SR * args.seconds) // HOP
    N = T * HOP
    torch.manual_seed(101)
    f0, control = torch.rand(B, 1, T, device=dev), torch.rand(B, 2, T, device=dev)
    k = max(5, min(args.steps, 30))
    with torch.no_grad():
        for _ in range(3):
            model(f0, control)
        per = timed_steps(k, lambda: model(f0, control))

        g = torch.Generator(device=dev).manual_seed(5000)
        n_sets = (160 << 20) // ((f0.numel() + control.numel()) * 4) + 1
        f0_sets = torch.rand(n_sets, B, 1, T, device=dev, generator=g)
        control_sets = torch.rand(n_sets, B, 2, T, device=dev, generator=g)
        thr_ms = lane_throughput_ms(model, dev, lambda i: (f0_sets[i], control_sets[i]), k, n_sets)
        del f0_sets, control_sets
        eng = model._engine_for(f0)
        eng.set_profiling(True)
        acc = {}
        for _ in range(5):
            flush.fill_(1)
            model(f0, control)
            for kk, v in eng.stage_times_ms().items():
                acc[kk] = acc.get(kk, 0.0) + v / 5
        eng.set_profiling(False)
    lat_ms = sum(per) / len(per)
    ms = thr_ms
    audio_ms = acc.get("audio_fused", 0.0)
    algo = B * T * ALGO_BYTES_PER_UTT_FRAME
    achieved = algo / (audio_ms * 1e-3) / 1e9 if audio_ms > 0 else None
    # arithmetic view: 1,701 sines per sample go through the SFU (16 lanes / SM / clock): the pipe that bounds this kernel
    sfu = None
    if audio_ms > 0 and clocks_now and clocks_now.get("sm_mhz"):
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        sfu_peak = n_sm * 16 * clocks_now["sm_mhz"] * 1e6
        sfu = {"sines_per_sample": 1701, "mufu_per_s": B * N * 1701 / (audio_ms * 1e-3), "mufu_peak_per_s": sfu_peak,
               "frac_of_sfu_peak": B * N * 1701 / (audio_ms * 1e-3) / sfu_peak}
    return {
        "workload": "NEWT MLP forward, batch %d x %g s @ 16 kHz (BASELINE.json configs[2])" % (B, args.seconds),
        "ms_per_step": ms, "steps": k, "value": B * N / (ms * 1e-3), "unit": "samples/s",
        "timing": "as the headline: `value` = %d steps, two in flight; `latency` = one forward at a time, L2 flushed" % k,
        "latency": {"ms_per_forward": lat_ms, "ms_p90": pctl(per, 0.9)},
        "rtf_per_utterance": (ms * 1e-3) / (B * args.seconds),
        "parity_max_abs_vs_golden": {"kat_randinit_newt": parity_max_abs(model, golden_case("kat_randinit_newt"), dev)},
        "roofline": {"kernel": "nws_audio_tc_kernel<MLP>", "bound": "hbm", "achieved": achieved, "peak": peaks_hbm, "unit": "GB/s",
                     "frac": (achieved / peaks_hbm) if achieved else None, "traffic": ncu_traffic("newt"),
                     "algorithmic_bytes_per_launch": algo, "kernel_ms": audio_ms,
                     "kernel_share_of_step": audio_ms / sum(acc.values()) if acc else None,
                     "issue_view": issue_view("newt", audio_ms, clocks_now), "sfu_view": sfu},
        "stages_ms": acc, "stages_order": "serial (stage profiling disables the pipelined order)",
    }


def config_c4(args, cpu_model, dev, flush):
    """BASELINE.json configs[3]: B = 1 at the buffer sizes of scripts/time_buffer_sizes.py — (a) the reference script's
    stateless forward (nothing carried between buffers, time_buffer_sizes.py:66-72), cold (L2 flushed before each
    call) and warm; (b) the stateful SynthStream.push of the same number of samples; (c) the CPU port beside them."""
    import copy
    import torch
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    out = {"workload": "B = 1 buffer sweep (BASELINE.json configs[3]; scripts/time_buffer_sizes.py:13,66-75)",
           "timing": "CUDA events per call on the launch stream; `cold` = 256 MiB L2 flush before every call, `warm` = back to back, "
                     "`graph_replay` = the forward captured once into a CUDA graph (no per-call host work ahead of the first kernel)",
           "sizes": BUFFER_SIZES, "stream_sizes": STREAM_SIZES}
    iters = 40
    for variant in ("fastnewt", "newt"):
        model = copy.deepcopy(cpu_model)
        if variant == "fastnewt":
            model.newt = FastNEWT(model.newt)
        model = model.to(dev)
        rows = {}
        with torch.no_grad():
            for bs in BUFFER_SIZES:
                T = bs // HOP
                torch.manual_seed(bs)
                f0, control = torch.rand(1, 1, T, device=dev), torch.rand(1, 2, T, device=dev)
                for _ in range(5):
                    model(f0, control)
                torch.cuda.synchronize(dev)

                def run(with_flush):
                    evs = []
                    for _ in range(iters):
                        if with_flush:
                            flush.fill_(1)
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        model(f0, control)
                        b.record()
                        evs.append((a, b))
                    torch.cuda.synchronize(dev)
                    return [a.elapsed_time(b) for a, b in evs]

                cold, warm = run(True), run(False)
                # the same forward captured once into a CUDA graph and replayed: device-side latency of the launch
                # chain without the per-call host work (Python, ctypes, torch.empty) ahead of the first kernel
                graph_ms = None
                try:
                    # (draws injected: reserving a range of torch's generator is host work that cannot be captured)
                    u_g, nz_g = torch.rand(101, device=dev), torch.rand(HOP * T - 1, device=dev)
                    model(f0, control, phase_shift=u_g, noise=nz_g)
                    torch.cuda.synchronize(dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        y_static = model(f0, control, phase_shift=u_g, noise=nz_g)
                    g.replay()
                    torch.cuda.synchronize(dev)
                    evs = []
                    for _ in range(iters):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        g.replay()
                        b.record()
                        evs.append((a, b))
                    torch.cuda.synchronize(dev)
                    graph_ms = statistics.median([a.elapsed_time(b) for a, b in evs])
                    del g, y_static
                except Exception as e:   # informative only
                    print("graph capture failed at bs %d: %s" % (bs, e), file=sys.stderr)
                rows[str(bs)] = {"ms_median_cold": statistics.median(cold), "ms_p90_cold": pctl(cold, 0.9),
                                 "ms_median_warm": statistics.median(warm), "ms_p90_warm": pctl(warm, 0.9),
                                 "ms_median_graph_replay": graph_ms,
                                 "rtf_warm": statistics.median(warm) * 1e-3 / (bs / SR)}
            # stateful streaming: pushes of bs samples (bs / 128 frames) into one running stream
            stream_rows = {}
            for bs in STREAM_SIZES:
                n = bs // HOP
                st = model.stream(batch_size=1, max_frames=max(n, 2))
                st.reset()
                torch.manual_seed(bs)
                f0, control = torch.rand(1, 1, n, device=dev) * 200 + 100, torch.rand(1, 2, n, device=dev)
                for _ in range(5):
                    st.push(f0, control)
                torch.cuda.synchronize(dev)
                evs = []
                for _ in range(iters):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    st.push(f0, control)
                    b.record()
                    evs.append((a, b))
                torch.cuda.synchronize(dev)
                ts = [a.elapsed_time(b) for a, b in evs]
                stream_rows[str(bs)] = {"ms_median": statistics.median(ts), "ms_p90": pctl(ts, 0.9),
                                        "rtf": statistics.median(ts) * 1e-3 / (bs / SR)}
                del st
        par = {}
        for bs in (256, 4096, 32768):
            c = golden_case("sweep_randinit", "bs%d_" % bs)
            par["bs%d" % bs] = parity_max_abs(model, c, dev, "out_fast" if variant == "fastnewt" else "out")
        out[variant] = {"stateless_forward": rows, "stream_push": stream_rows, "parity_max_abs_vs_golden": par}
    return out


def config_c4_cpu(cpu_model, threads=8):
    """The CPU port at the same buffer sizes (rank 0, N = 1 only; a bounded sample: 3 calls per size after one warm-up)."""
    import torch
    from oracle import nws_oracle as oracle
    torch.set_num_threads(min(threads, os.cpu_count() or 1))
    w = {k: v.detach().cpu().clone() for k, v in cpu_model.state_dict().items()}
    lut = oracle.build_lookup_table(w)
    out = {"kind": "port", "cores": torch.get_num_threads(), "sample": "3 calls per size after 1 warm-up, B = 1, oracle port"}
    for variant, table in (("fastnewt", lut), ("newt", None)):
        rows = {}
        for bs in BUFFER_SIZES:
            T = bs // HOP
            torch.manual_seed(bs)
            f0, control = torch.rand(1, 1, T), torch.rand(1, 2, T)
            ts = []
            for i in range(4):
                t0 = time.perf_counter()
                u, noise = oracle.draw_rng(T)
                oracle.forward(w, f0, control, u, noise, lut=table, faithful_loop=True)
                if i:
                    ts.append((time.perf_counter() - t0) * 1e3)
            rows[str(bs)] = {"ms_median": statistics.median(ts), "rtf": statistics.median(ts) * 1e-3 / (bs / SR)}
        out[variant] = rows
    return out


def config_c5(args, model, dev, rank, world, dist):
    """BASELINE.json configs[4]: one GLOBAL batch of 2048 utterances x 4 s, split over the ranks in contiguous slices
    (sharding.shard_bounds: 256 per GPU at N = 8), forward on every rank, then the audio gathered on every rank over
    NCCL (sharding.gather_audio) — strong scaling of a fixed job.  Time = max over ranks of (forward + gather)."""
    import torch
    from neural_waveshaping_synthesis_b200.sharding import aggregate_throughput, gather_audio, shard_bounds
    G, T = args.c5_global_batch, int(SR * args.seconds) // HOP
    N = T * HOP
    lo, hi = shard_bounds(G, rank, world)
    gen = torch.Generator().manual_seed(2048)            # every rank draws the same global inputs and takes its slice
    f0_all, control_all = torch.rand(G, 1, T, generator=gen), torch.rand(G, 2, T, generator=gen)
    u = torch.rand(101, generator=gen).to(dev)
    noise = torch.rand(N - 1, generator=gen).to(dev)
    f0, control = f0_all[lo:hi].contiguous().to(dev), control_all[lo:hi].contiguous().to(dev)
    k = 5
    fwd, gat = [], []
    with torch.no_grad():
        for i in range(2 + k):
            torch.cuda.synchronize(dev)
            if dist is not None:
                dist.barrier()
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            y = model(f0, control, phase_shift=u, noise=noise)
            b.record()
            full = gather_audio(y, G)
            c.record()
            torch.cuda.synchronize(dev)
            if i >= 2:
                fwd.append(a.elapsed_time(b))
                gat.append(b.elapsed_time(c))
        # parity: rows of the gathered batch that OTHER ranks rendered (first and last utterance of the job) against the
        # same utterances rendered alone on this rank with the same draws (fp32 round-off: the reverb pairs utterances)
        par = {}
        for j in sorted({0, G // 2, G - 1}):
            solo = model(f0_all[j:j + 1].to(dev), control_all[j:j + 1].to(dev), phase_shift=u, noise=noise)
            par["row%d_vs_solo_max_abs" % j] = float((full[j] - solo[0]).abs().max())
        shape_ok = tuple(full.shape) == (G, N)
        del full
    f_ms, g_ms = sum(fwd) / k, sum(gat) / k
    tot_max, _ = aggregate_throughput(f_ms + g_ms, 0.0, dev)
    f_max, _ = aggregate_throughput(f_ms, 0.0, dev)
    g_max, _ = aggregate_throughput(g_ms, 0.0, dev)
    res = {"workload": "%s forward of a global batch of %d utterances x %g s sharded over %d GPU(s) + NCCL gather of the audio "
                       "(BASELINE.json configs[4])" % ("FastNEWT" if args.variant == "fastnewt" else "NEWT", G, args.seconds, world),
           "global_batch": G, "utterances_per_gpu": hi - lo, "n_gpus": world, "scaling": "strong", "steps": k,
           "forward_ms_max": f_max, "gather_ms_max": g_max, "ms_per_job_max": tot_max,
           "gather": ("ncclAllGather of [%d, %d] fp32 per rank straight into the full batch (%.1f MB in, %.1f MB out per rank)" %
                      (hi - lo, N, (hi - lo) * N * 4 / 1e6, G * N * 4 / 1e6)) if world > 1 else "single rank: nothing to gather",
           "value": G * N / (tot_max * 1e-3), "value_forward_only": G * N / (f_max * 1e-3), "unit": "samples/s",
           "gathered_shape_ok": shape_ok, "parity": par}
    # the same job in two waves per rank, the gather of wave 0 overlapping the forward of wave 1
    # (sharding.forward_and_gather: wave w = rows [w G/2, (w+1) G/2) split over the ranks)
    waves = 2
    if world > 1 and G % (world * waves) == 0:
        from neural_waveshaping_synthesis_b200.sharding import forward_and_gather, wave_bounds
        rows = torch.cat([torch.arange(*wave_bounds(G, rank, world, waves, w)) for w in range(waves)])
        f0_w, control_w = f0_all[rows].contiguous().to(dev), control_all[rows].contiguous().to(dev)
        ov = []
        with torch.no_grad():
            for i in range(2 + k):
                torch.cuda.synchronize(dev)
                dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                full = forward_and_gather(model, f0_w, control_w, G, waves=waves, phase_shift=u, noise=noise)
                b.record()
                torch.cuda.synchronize(dev)
                if i >= 2:
                    ov.append(a.elapsed_time(b))
            par_w = {}
            for j in sorted({0, G // 2, G - 1}):
                solo = model(f0_all[j:j + 1].to(dev), control_all[j:j + 1].to(dev), phase_shift=u, noise=noise)
                par_w["row%d_vs_solo_max_abs" % j] = float((full[j] - solo[0]).abs().max())
            ok_w = tuple(full.shape) == (G, N)
            del full
        ov_max, _ = aggregate_throughput(sum(ov) / k, 0.0, dev)
        res["overlapped"] = {"protocol": "%d waves per rank; ncclAllGather of wave w (async, into its rows of the full batch) "
                                         "beside the forward of wave w+1" % waves,
                             "ms_per_job_max": ov_max, "value": G * N / (ov_max * 1e-3), "unit": "samples/s",
                             "gathered_shape_ok": ok_w, "parity": par_w}
        if ov_max < tot_max:
            res["value_serial_protocol"] = res["value"]
            res["value"] = res["overlapped"]["value"]
            res["ms_per_job_max_serial_protocol"], res["ms_per_job_max"] = tot_max, ov_max
    return res


def workload_config(args):
    return {"workload": "%s forward, batch %d x %g s @ 16 kHz per GPU (BASELINE.json configs[%d])" %
            ("FastNEWT LUT" if args.variant == "fastnewt" else "NEWT MLP", args.batch_per_gpu, args.seconds,
             1 if args.variant == "fastnewt" else 2),
            "variant": args.variant, "batch_per_gpu": args.batch_per_gpu, "global_batch": args.batch_per_gpu * args.gpus,
            "seconds": args.seconds, "frames": int(SR * args.seconds) // HOP, "sample_rate": SR,
            "inputs": ("torch.rand f0/control as scripts/time_forward_pass.py:27-40; random-init newt.gin weights, seed 0"
                       if getattr(args, "inputs", "rand") == "rand" else
                       "violin checkpoint (tests/golden/weights_vn.npz); f0 = 440 Hz x 2^(0.5 sin) vibrato scaled by "
                       "U[0.25,1.5) per utterance, loudness LFO, control normalised with the checkpoint's data_mean/std"),
            "rng": "on-device Philox draws inside the timed region",
            "l2": "inputs larger than L2: the steps rotate through 160 MiB of distinct device-resident input batches (no flush: "
                  "%d forwards are in flight, on %d streams / engines)" % (LANES, LANES),
            "parallelism": "dp%d (utterance shards, no data-path collective)" % args.gpus}


def run_b200(args):
    import torch
    rank, local_rank, world = env_rank()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    affinity = None
    if world > 1:
        # one process per GPU on one host: give every rank its own slice of the cores this job may use, so eight
        # host threads landing 16 MB per millisecond do not migrate over (and evict each other from) the same cores
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            os.sched_setaffinity(0, mine)
            affinity = "%d of %d cores (%d..%d)" % (len(mine), len(cores), mine[0], mine[-1])
        except Exception as e:   # not fatal: the numbers are still valid, only noisier
            affinity = "unchanged (%s)" % e

    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    from neural_waveshaping_synthesis_b200 import _lib
    import copy
    cpu_model = build_weights()
    if args.inputs == "realistic":            # the violin checkpoint instead of the random-init weights
        import numpy as np
        zw = np.load(os.path.join(REPO, "tests", "golden", "weights_vn.npz"))
        cpu_model.load_state_dict({k: torch.from_numpy(zw[k]) for k in zw.files if not k.startswith("data_")})
    model = copy.deepcopy(cpu_model)          # .to() moves a module in place: keep the CPU copy apart
    if args.variant == "fastnewt":
        model.newt = FastNEWT(model.newt)
    model = model.to(dev)
    B, T = args.batch_per_gpu, int(SR * args.seconds) // HOP
    N = T * HOP
    torch.manual_seed(1 + rank)
    if args.inputs == "realistic":
        import math
        import numpy as np
        z = np.load(os.path.join(REPO, "tests", "golden", "weights_vn.npz"))
        u = torch.linspace(0, 1, T)
        f0 = 440.0 * torch.pow(2.0, 0.5 * torch.sin(2 * math.pi * 1.5 * u)) * (0.25 + 1.25 * torch.rand(B, 1, 1))
        loud = (0.10 + 0.03 * torch.sin(2 * math.pi * 3 * u)).view(1, 1, T).expand(B, 1, T)
        mean, std = z["data_mean"], z["data_std"]
        f0_host = f0.float().contiguous().pin_memory()
        control_host = torch.cat(((f0 - float(mean[0, 0])) / float(std[0, 0]), (loud - float(mean[1, 0])) / float(std[1, 0])),
                                 dim=1).float().contiguous().pin_memory()
    else:
        f0_host = torch.rand(B, 1, T).pin_memory()
        control_host = torch.rand(B, 2, T).pin_memory()
    f0, control = f0_host.to(dev), control_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load_library()

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed_steps(k, fn, with_flush=True):
        evs = []
        for _ in range(k):
            if with_flush:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in evs]

    with torch.no_grad():
        for _ in range(args.warmup):
            model(f0, control)
        barrier()
        # clocks are sampled on rank 0 only (every GPU of the box runs the same load; eight nvidia-smi loops at 20 ms
        # were part of the host-side noise of the 8-GPU runs) and over ALL timed regions of this process
        sampler = ClockSampler(dev.index) if rank == 0 else None
        if sampler:
            sampler.start()
        per_step = timed_steps(args.steps, lambda: model(f0, control))
        barrier()
        lat_ms = sum(per_step) / len(per_step)          # one forward at a time, L2 flushed before each: the latency view
        step_p90 = sorted(per_step)[min(len(per_step) - 1, int(0.9 * len(per_step)))]   # mean + p90
        # ---- the timed region of `value`: K steps, two in flight
        n_sets = (160 << 20) // ((f0.numel() + control.numel()) * 4) + 1     # > the 126 MB L2

        if args.inputs == "rand":   # two launches for all the sets (not two per set: they would crowd the launch list)
            g = torch.Generator(device=dev).manual_seed(1000 * (1 + rank))
            f0_sets = torch.rand(n_sets, B, 1, T, device=dev, generator=g)
            control_sets = torch.rand(n_sets, B, 2, T, device=dev, generator=g)

        def make_inputs(i):
            return (f0_sets[i], control_sets[i]) if args.inputs == "rand" else (f0, control)
        barrier()
        lib.nws_launch_count(1)
        step_ms = lane_throughput_ms(model, dev, make_inputs, args.steps, n_sets if args.inputs == "rand" else 1)
        launches = int(lib.nws_launch_count(0))
        barrier()

        # ---- dominant-kernel time (library stage events) on the same workload
        eng = model._engine_for(f0)
        eng.set_profiling(True)
        stage_acc = {}
        for _ in range(args.steps):
            flush.fill_(1)
            model(f0, control)
            for k, v in eng.stage_times_ms().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v / args.steps
        eng.set_profiling(False)

        # ---- the library's alternatives on the same workload (A/B, same timing as `value`): the fp32 SIMT recurrence
        # instead of the tensor-core one, and the FIR noise branch inside the fused audio kernel instead of its own launch
        variants_ms = {"default": lat_ms}
        k_ab = max(5, min(args.steps, 30))
        for name, on, off in (("gru_fp32_simt", lambda: eng.set_gru_impl(0), lambda: eng.set_gru_impl(1)),
                              ("noise_branch_in_audio_kernel", lambda: eng.set_noise_fused(True), lambda: eng.set_noise_fused(False))):
            on()
            for _ in range(3):
                model(f0, control)
            per = timed_steps(k_ab, lambda: model(f0, control))
            off()
            variants_ms[name] = sum(per) / len(per)

        # ---- end to end through the public API for host-resident batches (streaming.HostPipeline, the loop of
        # scripts/resynthesise_dataset.py): every step copies its inputs from pinned host memory, runs the forward
        # and reads the audio back to pinned host memory; upload of step i+1 and download of step i-1 overlap the
        # forward of step i.  Every byte of every step is inside the timed region, which ends when the last
        # result has landed on the host.
        from neural_waveshaping_synthesis_b200.streaming import HostPipeline

        pipe = HostPipeline(model, dev)     # one pipeline: its staging buffers (device inputs, two pinned 16 MB result
                                            # buffers) are allocated by the warm-up pass, not inside the timed region

        def e2e_run(n):
            landed = 0
            for _, audio in pipe.run((f0_host, control_host) for _ in range(n)):
                landed += 1
            assert landed == n and audio.shape == (B, N)

        e2e_run(max(args.warmup, 3))
        barrier()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        t0 = time.perf_counter()
        e2e_run(args.steps)
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) / args.steps
        h2d_per_step, d2h_per_step = pipe.h2d_bytes // args.steps, pipe.d2h_bytes // args.steps
        clocks = sampler.stop() if sampler else None   # sampled over the three timed regions above (device-timed, stages, e2e)

        # ---- fp32 FMA issue-rate probe (the compute roofline's denominator, measured on this box)
        ffma_tflops = None
        try:
            import ctypes
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            probe_out = torch.empty(n_sm * 8 * 256, dtype=torch.float32, device=dev)
            flops = ctypes.c_double(0.0)
            stream = torch.cuda.current_stream(dev).cuda_stream
            best = None
            for i in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _lib.check(lib.nws_selftest_ffma_peak(probe_out.data_ptr(), n_sm * 8, 1024, ctypes.byref(flops), stream))
                b.record()
                torch.cuda.synchronize(dev)
                if i > 0:
                    best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
            ffma_tflops = flops.value / (best * 1e-3) / 1e12
        except Exception as e:   # the probe is informative only
            print("ffma probe failed: %s" % e, file=sys.stderr)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")

    # ---- the other BASELINE configurations, inside the same line
    configs = {}
    if not args.no_configs:
        clocks_mid = clocks if clocks and clocks.get("sm_mhz") else {"sm_mhz": peaks.get("sm_max_mhz", 1965.0)}
        configs["c2"] = {"workload": "= this line's headline (`value`, `e2e`, `roofline`)",
                         "parity_max_abs_vs_golden": {"kat_randinit_fast": parity_max_abs(model, golden_case("kat_randinit_fast"), dev)}
                         if args.variant == "fastnewt" and args.inputs == "rand" else None}
        if world == 1:
            configs["c3"] = config_c3(args, cpu_model, dev, timed_steps, flush, hbm_peak, clocks_mid)
            configs["c4"] = config_c4(args, cpu_model, dev, flush)
        barrier()
        configs["c5"] = config_c5(args, model, dev, rank, world, dist)
        barrier()

    # ---- aggregate over ranks: time = max over ranks, samples = sum
    from neural_waveshaping_synthesis_b200.sharding import aggregate_throughput
    step_ms_max, total_samples = aggregate_throughput(step_ms, float(B * N), dev)
    e2e_ms_max, _ = aggregate_throughput(e2e_s * 1e3, float(B * N), dev)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    audio_ms = stage_acc.get("audio_fused", 0.0)
    algo_bytes = B * T * ALGO_BYTES_PER_UTT_FRAME          # 262,000 B per 4 s utterance
    achieved = algo_bytes / (audio_ms * 1e-3) / 1e9 if audio_ms > 0 else None
    line = {
        "metric": "audio samples/sec", "value": total_samples / (step_ms_max * 1e-3), "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "latency": {"ms_per_forward": lat_ms, "ms_p90": step_p90, "steps": len(per_step),
                    "how": "one forward at a time on rank 0, CUDA events around each, 256 MiB L2 flush before each "
                           "(round 1's `ms_per_step`; `value` is the throughput of the same forwards two in flight)"},
        "rtf_per_utterance": (step_ms_max * 1e-3) / (B * args.seconds),
        "rtf_batch": (step_ms_max * 1e-3) / args.seconds,
        "config": workload_config(args),
        "e2e": {"value": total_samples / (e2e_ms_max * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms_max,
                "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": d2h_per_step},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "nws_audio_tc_kernel<%s>" % ("LUT" if args.variant == "fastnewt" else "MLP"),
                     "bound": "hbm", "binding_limit": "warp-instruction issue (see issue_view): the path is not HBM-bound by construction",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": (achieved / hbm_peak) if achieved else None, "traffic": ncu_traffic(args.variant),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": audio_ms,
                     "kernel_share_of_step": audio_ms / sum(stage_acc.values()) if stage_acc else None,
                     "issue_view": issue_view(args.variant, audio_ms, clocks, ffma_tflops)},
        "stages_ms": stage_acc,
        "stages_order": "serial (stage profiling disables the pipelined order; `ms_per_step` is the pipelined forward)",
        "recurrence": ("tensor cores (mma.sync m16n8k16, fp16-split operands, 8 utterances per CTA: csrc/nws_gru_mma.cu)"
                       if B >= 64 else "fp32 SIMT, one utterance per CTA (fewer than 64 utterances)"),
        "variants_ms_per_step_rank0": variants_ms,
        "host_affinity": affinity,
    }
    if configs:
        line["configs"] = configs
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only
        line["cpu_baseline"] = best_cpu_baseline(cpu_model, args, T)
        if configs.get("c4") is not None:
            configs["c4"]["cpu_baseline"] = config_c4_cpu(cpu_model)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    # communicator creation when NCCL_DEBUG is set), so file descriptor 1 points at stderr while the benchmark
    # runs and the JSON line goes to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
    for ln in lines[:-1]:
        print(ln, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()
