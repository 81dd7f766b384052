mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 400 gpurun_out/bench_n8.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_n8.json').read())
print(d['value'], d['ms_per_step'], d['latency'], d['e2e'], d['host_affinity'], d['clocks'])
c = d['configs']['c5']
print({k: v for k, v in c.items() if k not in ('workload', 'gather', 'overlapped', 'parity')}, c.get('overlapped'))
P
