mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 300 gpurun_out/bench_fast.err; cat gpurun_out/bench_fast.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'], d['stages_ms'])"
