mkdir -p gpurun_out
timeout 300 python bench.py --variant newt --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_newt.json 2> gpurun_out/b_newt.err; tail -c 200 gpurun_out/b_newt.err
timeout 300 python bench.py --inputs realistic --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_real.json 2> gpurun_out/b_real.err; tail -c 200 gpurun_out/b_real.err
python - <<'P'
import json
for f in ("b_newt", "b_real"):
    d = json.loads(open("gpurun_out/%s.json" % f).read())
    print(f, d["ms_per_step"], d["latency"]["ms_per_forward"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["variant"])
P
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
