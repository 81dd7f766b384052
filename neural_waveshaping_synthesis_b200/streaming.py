"""Batches that live in host memory: upload, forward and download as a three-stage software pipeline.

The reference's `scripts/resynthesise_dataset.py:14-25` moves each batch to the device, runs the model and
brings the audio back, one after the other.  On a B200 the forward of a 64 x 4 s batch takes ~1.4 ms while its
16 MB of audio needs ~0.3 ms of PCIe time, so the copies are worth hiding: `HostPipeline` uploads batch i+1
and downloads batch i-1 on their own streams while batch i computes.  The upload is issued *ahead* of the
download of the previous result — a small H2D queued behind a 16 MB D2H otherwise delays the next forward.
"""
import torch


class HostPipeline:
    """`for meta, audio in HostPipeline(model, device).run(batches)`; `batches` yields `(f0, control)` or
    `(f0, control, meta)` with host tensors (pinned memory if the copies are to be asynchronous).  `audio` is
    a pinned host tensor [B, 128*T], valid until the next iteration of the generator."""

    def __init__(self, model, device):
        self.model = model
        self.device = torch.device(device)
        self.up_stream = torch.cuda.Stream(self.device)
        self.down_stream = torch.cuda.Stream(self.device)
        self._dev_in = [None, None]
        self._host_out = [None, None]
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
        main = torch.cuda.current_stream(self.device)
        try:
            staged = self._upload(next(it), 0, None)
        except StopIteration:
            return
        done_prev = None          # end of forward i-1 (the last reader of slot (i+1)&1)
        pending = None            # (event, meta, host tensor) of step i-1
        i = 0
        with torch.no_grad():
            while staged is not None:
                (f0, control), ready, meta = staged
                main.wait_event(ready)
                y = self.model(f0, control)
                done = torch.cuda.Event()
                done.record(main)
                try:                                   # stage the next batch before this result's download
                    staged = self._upload(next(it), (i + 1) & 1, done_prev)
                except StopIteration:
                    staged = None
                slot = i & 1
                host = self._host_out[slot]
                if host is None or host.shape != y.shape:
                    host = torch.empty(y.shape, dtype=torch.float32).pin_memory()
                    self._host_out[slot] = host
                with torch.cuda.stream(self.down_stream):
                    self.down_stream.wait_event(done)
                    host.copy_(y, non_blocking=True)
                    y.record_stream(self.down_stream)
                    landed = torch.cuda.Event()
                    landed.record(self.down_stream)
                self.d2h_bytes += y.numel() * 4
                if pending is not None:
                    pending[0].synchronize()
                    yield pending[1], pending[2]
                pending = (landed, meta, host)
                done_prev = done
                i += 1
        if pending is not None:
            pending[0].synchronize()
            yield pending[1], pending[2]
