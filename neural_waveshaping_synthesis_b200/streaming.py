"""Batches that live in host memory: upload, forward and download as a three-stage software pipeline.

The reference's `scripts/resynthesise_dataset.py:14-25` moves each batch to the device, runs the model and
brings the audio back, one after the other.  On a B200 the forward of a 64 x 4 s batch takes ~1.4 ms while its
16 MB of audio needs ~0.3 ms of PCIe time, so the copies are worth hiding: `HostPipeline` uploads batch i+1
and downloads batch i-1 on their own streams while batch i computes.  The upload is issued *ahead* of the
download of the previous result — a small H2D queued behind a 16 MB D2H otherwise delays the next forward.
"""
import ctypes

import torch

from . import _lib

HOP = 128


class HostPipeline:
    """`for meta, audio in HostPipeline(model, device).run(batches)`; `batches` yields `(f0, control)` or
    `(f0, control, meta)` with host tensors (pinned memory if the copies are to be asynchronous).  `audio` is
    a pinned host tensor [B, 128*T], valid until the next iteration of the generator.

    Batch i is uploaded on one stream, rendered on compute stream i % lanes by engine i % lanes, and downloaded on a third
    stream.  With `lanes` = 2 (default) two forwards are in flight: a forward starts with its encoder alone on a few SMs
    and ends with a reverb tail, and at 64 utterances those edges are a quarter of its latency (0.98 ms against 0.73 ms
    of work at the chip's large-batch rate) — the neighbour batch fills them.  Results are yielded in order."""

    def __init__(self, model, device, lanes: int = 2):
        self.model = model
        self.device = torch.device(device)
        self.lanes = max(1, int(lanes))
        self.depth = 2 * self.lanes          # staging slots (device inputs / device outputs / pinned results)
        self.up_stream = torch.cuda.Stream(self.device)
        self.down_stream = torch.cuda.Stream(self.device)
        self.compute = [torch.cuda.Stream(self.device) for _ in range(self.lanes)] if self.lanes > 1 else None
        self._dev_in = [None] * self.depth
        self._dev_out = [None] * self.depth
        self._host_out = [None] * self.depth
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _upload(self, item, slot, after):
        f0, control = item[0], item[1]
        meta = item[2] if len(item) > 2 else None
        buf = self._dev_in[slot]
        if buf is None or buf[0].shape != f0.shape or buf[1].shape != control.shape:
            buf = (torch.empty(f0.shape, dtype=torch.float32, device=self.device),
                   torch.empty(control.shape, dtype=torch.float32, device=self.device))
            self._dev_in[slot] = buf
        with torch.cuda.stream(self.up_stream):
            if after is not None:
                self.up_stream.wait_event(after)       # the forward that last read this slot
            buf[0].copy_(f0, non_blocking=True)
            buf[1].copy_(control, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.up_stream)
        self.h2d_bytes += f0.numel() * 4 + control.numel() * 4
        return buf, ready, meta

    def run(self, batches):
        it = iter(batches)
        caller = torch.cuda.current_stream(self.device)
        D = self.depth
        try:
            staged = self._upload(next(it), 0, None)
        except StopIteration:
            return
        done = [None] * D         # end of the forward that used staging slot s
        landed = [None] * D       # end of the download out of staging slot s
        pending = []              # (landed event, meta, host tensor) in batch order
        i = 0
        with torch.no_grad():
            while staged is not None:
                (f0, control), ready, meta = staged
                slot, lane = i % D, i % self.lanes
                cs = self.compute[lane] if self.compute else caller
                T = f0.shape[-1]
                y = self._dev_out[slot]
                if y is None or y.shape != (f0.shape[0], T * HOP):
                    y = torch.empty(f0.shape[0], T * HOP, dtype=torch.float32, device=self.device)
                    self._dev_out[slot] = y
                with torch.cuda.stream(cs):
                    cs.wait_event(ready)
                    if landed[slot] is not None:
                        cs.wait_event(landed[slot])    # the download that last read this output buffer
                    self.model.forward_lane(lane, f0, control, out=y)
                    d = torch.cuda.Event()
                    d.record(cs)
                done[slot] = d
                try:                                   # stage the next batch before this result's download
                    nxt = (i + 1) % D
                    staged = self._upload(next(it), nxt, done[nxt])
                except StopIteration:
                    staged = None
                host = self._host_out[slot]
                if host is None or host.shape != y.shape:
                    host = torch.empty(y.shape, dtype=torch.float32).pin_memory()
                    self._host_out[slot] = host
                with torch.cuda.stream(self.down_stream):
                    self.down_stream.wait_event(d)
                    host.copy_(y, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.down_stream)
                landed[slot] = ev
                self.d2h_bytes += y.numel() * 4
                pending.append((ev, meta, host))
                # keep `lanes` batches in flight; a yielded host tensor is not rewritten before D - lanes further batches
                if len(pending) > self.lanes:
                    ev0, m0, h0 = pending.pop(0)
                    ev0.synchronize()
                    yield m0, h0
                i += 1
        for ev0, m0, h0 in pending:
            ev0.synchronize()
            yield m0, h0
        if self.compute:
            for cs in self.compute:                    # leave the caller's stream ordered after everything issued here
                caller.wait_stream(cs)


class SynthStream:
    """Stateful synthesis of B parallel control streams, a few frames per call (`nws_stream_*`,
    include/nws_b200.h): GRU state, phase, interpolation neighbours, noise overlap and reverb history are kept
    on the device between pushes.

        stream = model.stream(batch_size=1, max_frames=8)      # model: NeuralWaveshaping on a CUDA device
        stream.reset()
        for f0_chunk, control_chunk in chunks:                   # [B,1,n], [B,C,n]
            audio = stream.push(f0_chunk, control_chunk)         # [B, 128*n]  (128*(n-1) for the first push)
        tail = stream.push(None, None, flush=True)               # the last hop

    The concatenated output equals the dry signal of one whole-utterance forward followed by the reverb as a
    causal convolution; it lags the input by one hop (hop h interpolates towards frame h+1)."""

    def __init__(self, model, batch_size: int, max_frames: int, device=None):
        self.model = model
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("SynthStream needs the model on a CUDA device (there is no CPU path)")
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.max_frames = int(batch_size), int(max_frames)
        self.engine = model._engine_for(torch.empty(0, device=self.device))
        self.lib = self.engine.lib
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stream_create(self.engine.handle, self.B, self.max_frames, ctypes.byref(h)))
        self.handle = h
        self._out = torch.empty(self.B, HOP * (self.max_frames + 1), dtype=torch.float32, device=self.device)

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.nws_stream_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, phase_shift=None):
        """Start of an utterance.  `phase_shift` ([101] uniform draw, test hook) as in NeuralWaveshaping.forward."""
        self.engine = self.model._engine_for(torch.empty(0, device=self.device))   # weights / LUT up to date
        up = None if phase_shift is None else phase_shift.detach().to(self.device, torch.float32).reshape(-1).contiguous()
        if up is not None and up.numel() != 101:
            raise ValueError("phase_shift must have 101 elements")
        gen = torch.cuda.default_generators[self.device.index]
        off = gen.get_offset()
        gen.set_offset(off + 4 * 64)
        self._seed, self._offset = gen.initial_seed() & 0xFFFFFFFFFFFFFFFF, off // 4
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stream_reset(self.handle, ctypes.c_void_p(0 if up is None else up.data_ptr()),
                                                 self._seed, self._offset, self._stream()))

    def window(self, n_frames: int):
        """(first_frame, window_frames) of the next push — what an injected `noise_window` must cover."""
        first, n = ctypes.c_longlong(), ctypes.c_int()
        _lib.check(self.lib.nws_stream_window(self.handle, int(n_frames), ctypes.byref(first), ctypes.byref(n)))
        return first.value, n.value

    def push(self, f0, control, flush: bool = False, noise_window=None, reverb: bool = True):
        from .models.modules.shaping import FastNEWT
        n = 0 if f0 is None else int(f0.shape[-1])
        if n:
            if f0.dim() != 3 or f0.shape[0] != self.B or f0.shape[1] != 1:
                raise ValueError("f0 must be [%d, 1, n] (got %s)" % (self.B, tuple(f0.shape)))
            if control.dim() != 3 or control.shape[0] != self.B or control.shape[1] < 2 or control.shape[2] != n:
                raise ValueError("control must be [%d, C>=2, %d] (got %s)" % (self.B, n, tuple(control.shape)))
            for name, t in (("f0", f0), ("control", control)):
                if t.dtype != torch.float32 or t.device != self.device:
                    raise ValueError("%s must be float32 on %s" % (name, self.device))
            f0, control = f0.contiguous(), control.contiguous()
        nz = None
        if noise_window is not None:
            nz = noise_window.detach().to(self.device, torch.float32).reshape(-1).contiguous()
            if nz.numel() != HOP * self.window(n)[1] - 1:
                raise ValueError("noise_window must have 128*window_frames-1 = %d elements" % (HOP * self.window(n)[1] - 1))
        n_out = ctypes.c_int()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stream_push(
                self.handle, ctypes.c_void_p(f0.data_ptr() if n else 0), ctypes.c_void_p(control.data_ptr() if n else 0),
                int(control.shape[1]) if n else 0, n, ctypes.c_void_p(0 if nz is None else nz.data_ptr()),
                1 if isinstance(self.model.newt, FastNEWT) else 0, 1 if flush else 0, 1 if reverb else 0,
                ctypes.c_void_p(self._out.data_ptr()), ctypes.byref(n_out), self._stream()))
        return self._dense(n_out.value)

    def _dense(self, n_out: int):
        # the library writes [B, 128*n_out] densely at the start of the buffer
        return self._out.reshape(-1)[: self.B * HOP * n_out].view(self.B, HOP * n_out).clone()
