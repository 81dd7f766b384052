timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 300 gpurun_out/bench_fast.err; cat gpurun_out/bench_fast.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'])"
timeout 300 python bench.py --steps 10 --warmup 3 --variant newt > gpurun_out/bench_newt.json 2> gpurun_out/bench_newt.err; cat gpurun_out/bench_newt.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'])"
timeout 300 python scripts/parity_report.py 2>&1 | tail -12
