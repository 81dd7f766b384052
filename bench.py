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

One JSON line on stdout (rank 0).  `value` is device-timed (CUDA events around each step on the
launch stream, inputs resident in HBM, L2 flushed between steps); `e2e` is the same forward through
the public module API from pinned host tensors with the H2D/D2H copies inside the timed region;
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
    ap.add_argument("--inputs", default="rand", choices=["rand", "realistic"],
                    help="rand = the reference timing scripts' torch.rand f0/control (the contract's workload); "
                         "realistic = violin checkpoint + vibrato around 110-660 Hz (SURVEY.md 8(d)'s second input set)")
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """nvidia-smi sampling of SM clock and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def ncu_traffic(variant):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/, written by scripts/ncu_summary.py); None if absent."""
    path = os.path.join(REPO, "profiles", "r1_ncu_audio_tc_%s.json" % ("lut" if variant == "fastnewt" else "mlp"))
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
    path = os.path.join(REPO, "profiles", "r1_ncu_audio_tc_%s.json" % ("lut" if variant == "fastnewt" else "mlp"))
    try:
        d = json.load(open(path))[0]
        g = lambda k: d[k]["value"]
        v = {"bound": "warp-instruction issue (fp32 SIMT epilogue + sine generation)",
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
            "rng": "on-device Philox draws inside the timed region", "l2": "flushed between timed steps (256 MiB write)",
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
        sampler = ClockSampler(dev.index)
        sampler.start()
        lib.nws_launch_count(1)
        per_step = timed_steps(args.steps, lambda: model(f0, control))
        launches = int(lib.nws_launch_count(0))
        barrier()
        clocks = sampler.stop()
        step_ms = sum(per_step) / len(per_step)
        step_p90 = sorted(per_step)[min(len(per_step) - 1, int(0.9 * len(per_step)))]   # mean + p90

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

        # ---- end to end through the public API for host-resident batches (streaming.HostPipeline, the loop of
        # scripts/resynthesise_dataset.py): every step copies its inputs from pinned host memory, runs the forward
        # and reads the audio back to pinned host memory; upload of step i+1 and download of step i-1 overlap the
        # forward of step i.  Every byte of every step is inside the timed region, which ends when the last
        # result has landed on the host.
        from neural_waveshaping_synthesis_b200.streaming import HostPipeline

        def e2e_run(n):
            pipe = HostPipeline(model, dev)
            landed = 0
            for _, audio in pipe.run((f0_host, control_host) for _ in range(n)):
                landed += 1
            assert landed == n and audio.shape == (B, N)
            return pipe

        e2e_run(args.warmup)
        barrier()
        t0 = time.perf_counter()
        pipe = e2e_run(args.steps)
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) / args.steps
        h2d_per_step, d2h_per_step = pipe.h2d_bytes // args.steps, pipe.d2h_bytes // args.steps

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

    # ---- aggregate over ranks: time = max over ranks, samples = sum
    from neural_waveshaping_synthesis_b200.sharding import aggregate_throughput
    step_ms_max, total_samples = aggregate_throughput(step_ms, float(B * N), dev)
    e2e_ms_max, _ = aggregate_throughput(e2e_s * 1e3, float(B * N), dev)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    audio_ms = stage_acc.get("audio_fused", 0.0)
    algo_bytes = B * T * ALGO_BYTES_PER_UTT_FRAME          # 262,000 B per 4 s utterance
    achieved = algo_bytes / (audio_ms * 1e-3) / 1e9 if audio_ms > 0 else None
    line = {
        "metric": "audio samples/sec", "value": total_samples / (step_ms_max * 1e-3), "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "ms_per_step_p90_rank0": step_p90,
        "rtf_per_utterance": (step_ms_max * 1e-3) / (B * args.seconds),
        "rtf_batch": (step_ms_max * 1e-3) / args.seconds,
        "config": workload_config(args),
        "e2e": {"value": total_samples / (e2e_ms_max * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms_max,
                "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": d2h_per_step},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "nws_audio_tc_kernel<%s>" % ("LUT" if args.variant == "fastnewt" else "MLP"),
                     "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": (achieved / hbm_peak) if achieved else None, "traffic": ncu_traffic(args.variant),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": audio_ms,
                     "kernel_share_of_step": audio_ms / sum(stage_acc.values()) if stage_acc else None,
                     "issue_view": issue_view(args.variant, audio_ms, clocks, ffma_tflops)},
        "stages_ms": stage_acc,
    }
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only
        line["cpu_baseline"] = best_cpu_baseline(cpu_model, args, T)
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
