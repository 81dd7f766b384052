mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 300 gpurun_out/bench_fast.err; cat gpurun_out/bench_fast.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'], d['stages_ms'])"
timeout 300 python bench.py --steps 10 --warmup 3 --variant newt --no-cpu-baseline > gpurun_out/bench_newt.json 2> gpurun_out/bench_newt.err; cat gpurun_out/bench_newt.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'])"
timeout 300 python scripts/parity_report.py --json gpurun_out/parity.json 2>&1 | tail -12
