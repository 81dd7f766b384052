"""Where the end-to-end step (bench.py `e2e`) spends its time beyond the device-timed forward: back-to-back
forwards with resident inputs, + H2D, + D2H (overlapped on a copy stream / serial), and the host enqueue time."""
import copy
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT  # noqa: E402

dev = torch.device("cuda", 0)
model = copy.deepcopy(bench.build_weights())
model.newt = FastNEWT(model.newt)
model = model.to(dev)
B, T = 64, 500
N = T * 128
f0_host = torch.rand(B, 1, T).pin_memory()
control_host = torch.rand(B, 2, T).pin_memory()
f0, control = f0_host.to(dev), control_host.to(dev)
out_host = [torch.empty(B, N).pin_memory() for _ in range(2)]
copy_stream = torch.cuda.Stream(dev)
STEPS = 40


def wall(fn, steps=STEPS, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3, t_host / steps * 1e3


copied = [None, None]


def step_overlap(i):
    y = model(f0_host.to(dev, non_blocking=True), control_host.to(dev, non_blocking=True))
    done = torch.cuda.Event()
    done.record()
    slot = i & 1
    if copied[slot] is not None:
        copied[slot].synchronize()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(done)
        out_host[slot].copy_(y, non_blocking=True)
        y.record_stream(copy_stream)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
    copied[slot] = ev


def step_overlap_resident(i):
    y = model(f0, control)
    done = torch.cuda.Event()
    done.record()
    slot = i & 1
    if copied[slot] is not None:
        copied[slot].synchronize()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(done)
        out_host[slot].copy_(y, non_blocking=True)
        y.record_stream(copy_stream)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
    copied[slot] = ev


up_stream = torch.cuda.Stream(dev)
dev_in = [(torch.empty_like(f0), torch.empty_like(control)) for _ in range(2)]
up_ev = [None, None]
prev_done = [None]


def upload(i):
    with torch.cuda.stream(up_stream):
        dev_in[i & 1][0].copy_(f0_host, non_blocking=True)
        dev_in[i & 1][1].copy_(control_host, non_blocking=True)
        e = torch.cuda.Event()
        e.record(up_stream)
    up_ev[i & 1] = e


def step_prefetch(i):
    # inputs of step i were uploaded while step i-1 computed; upload i+1 now
    if up_ev[i & 1] is None:
        upload(i)
    torch.cuda.current_stream().wait_event(up_ev[i & 1])
    a, b = dev_in[i & 1]
    y = model(a, b)
    done = torch.cuda.Event()
    done.record()
    slot = i & 1
    if copied[slot] is not None:
        copied[slot].synchronize()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(done)
        out_host[slot].copy_(y, non_blocking=True)
        y.record_stream(copy_stream)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
    copied[slot] = ev
    if prev_done[0] is not None:
        up_stream.wait_event(prev_done[0])   # buffer (i+1)&1 was last read by step i-1
    upload(i + 1)
    prev_done[0] = done


def step_serial(i):
    y = model(f0_host.to(dev, non_blocking=True), control_host.to(dev, non_blocking=True))
    out_host[i & 1].copy_(y, non_blocking=True)


with torch.no_grad():
    res = {}
    res["resident, back to back"] = wall(lambda i: model(f0, control))
    res["+ H2D"] = wall(lambda i: model(f0_host.to(dev, non_blocking=True), control_host.to(dev, non_blocking=True)))
    res["+ H2D + D2H on copy stream (bench e2e)"] = wall(step_overlap)
    res["+ H2D + D2H same stream"] = wall(step_serial)
    res["resident + D2H on copy stream"] = wall(step_overlap_resident)
    res["prefetched H2D + D2H on copy stream"] = wall(step_prefetch)
    eng = model._engine_for(f0)
    eng.set_pipeline(False)
    res["serial forward, resident"] = wall(lambda i: model(f0, control))
    eng.set_pipeline(True)
    # D2H alone
    y = model(f0, control)
    res["D2H alone"] = wall(lambda i: out_host[i & 1].copy_(y, non_blocking=True))
    for k, (ms, host) in res.items():
        print("%-45s %.3f ms/step   (host enqueue %.3f ms/step)" % (k, ms, host), flush=True)
