mkdir -p gpurun_out
for inp in rand realistic; do for ch in 1 2 4 8; do
NWS_TILE_CHUNK=$ch timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --inputs $inp > gpurun_out/b.json 2> gpurun_out/b.err; tail -c 200 gpurun_out/b.err
python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read()); print('$inp chunk $ch', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), {k:round(v,4) for k,v in d['stages_ms'].items() if v>0.03})"
done; done
