# Quick GPU check used while iterating (gpurun -- 'bash scripts/_run.sh'): parity suite + a short bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 600 gpurun_out/bench_fast.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_fast.json').read())
print(d['ms_per_step'], d['latency'], d['e2e'], d['roofline']['kernel_ms'], d['stages_ms'], d['clocks'], d['variants_ms_per_step_rank0'], d['gpu_launches'])
c = d.get('configs', {})
if 'c3' in c: print('c3', c['c3']['ms_per_step'], c['c3']['latency'], c['c3']['roofline']['kernel_ms'], c['c3']['parity_max_abs_vs_golden'], c['c3']['roofline'].get('sfu_view'))
if 'c4' in c:
    for v in ('fastnewt', 'newt'):
        print('c4', v, {k: (round(r['ms_median_warm'], 4), round(r['ms_median_graph_replay'] or 0, 4)) for k, r in c['c4'][v]['stateless_forward'].items()}, {k: round(r['ms_median'], 4) for k, r in c['c4'][v]['stream_push'].items()}, c['c4'][v]['parity_max_abs_vs_golden'])
if 'c5' in c: print('c5', c['c5'])
P
