mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 400 gpurun_out/bench_n2.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_n2.json').read())
print(d['value'], d['ms_per_step'], d['e2e'], d['host_affinity'], d['clocks'], d['variants_ms_per_step_rank0'])
c = d['configs']['c5']
print({k: v for k, v in c.items() if k not in ('workload', 'gather', 'overlapped', 'parity')}, c.get('overlapped'))
P
PYTHONPATH=. timeout 300 python - <<'P'
# forward time per batch size on one GPU (pipelined order, FastNEWT): the large-batch regime
import torch, bench
from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
m = bench.build_weights(); m.newt = FastNEWT(m.newt); m = m.to('cuda:0')
for B in (64, 96, 128, 192, 256, 512, 1024):
    f0, c = torch.rand(B, 1, 500, device='cuda'), torch.rand(B, 2, 500, device='cuda')
    with torch.no_grad():
        for _ in range(3): m(f0, c)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): m(f0, c)
        b.record(); torch.cuda.synchronize()
    print("B %4d forward %.3f ms = %.2f us per utterance" % (B, a.elapsed_time(b) / 5, a.elapsed_time(b) / 5 / B * 1e3), flush=True)
P
