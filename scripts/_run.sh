# Quick GPU check used while iterating (gpurun -- 'bash scripts/_run.sh'): parity suite + a short bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 300 gpurun_out/bench_fast.err
python -c "import json; d=json.loads(open('gpurun_out/bench_fast.json').read()); print(d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'], d['stages_ms'])"
