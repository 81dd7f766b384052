mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/b.json 2> gpurun_out/b.err; tail -c 300 gpurun_out/b.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/b.json').read())
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['variants_ms_per_step_rank0'])
P
PYTHONPATH=. timeout 300 python - <<'P'
import time, torch, bench
from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
from neural_waveshaping_synthesis_b200.streaming import HostPipeline
m = bench.build_weights(); m.newt = FastNEWT(m.newt); m = m.to('cuda:0')
f0h, ch = torch.rand(64, 1, 500).pin_memory(), torch.rand(64, 2, 500).pin_memory()
for lanes in (1, 2, 3):
    pipe = HostPipeline(m, 'cuda:0', lanes=lanes)
    for _ in pipe.run((f0h, ch) for _ in range(6)): pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    for _ in pipe.run((f0h, ch) for _ in range(n)): pass
    torch.cuda.synchronize()
    print("lanes %d: e2e %.4f ms per batch" % (lanes, (time.perf_counter() - t0) / n * 1e3), flush=True)
# device-resident two-lane throughput
f0, c = f0h.cuda(), ch.cuda()
streams = [torch.cuda.Stream() for _ in range(2)]
outs = [torch.empty(64, 64000, device='cuda') for _ in range(2)]
with torch.no_grad():
    for k in range(6):
        with torch.cuda.stream(streams[k % 2]): m._forward_lane(k % 2, f0, c, out=outs[k % 2])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s_ in streams: s_.wait_event(a)
    n = 200
    for k in range(n):
        with torch.cuda.stream(streams[k % 2]): m._forward_lane(k % 2, f0, c, out=outs[k % 2])
    for s_ in streams: torch.cuda.current_stream().wait_stream(s_)
    b.record(); torch.cuda.synchronize()
print("device-resident, two lanes: %.4f ms per batch" % (a.elapsed_time(b) / n))
P
