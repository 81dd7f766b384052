timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 300 gpurun_out/bench_fast.err; cat gpurun_out/bench_fast.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
