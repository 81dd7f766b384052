mkdir -p gpurun_out
for inp in rand realistic; do for ch in 8 16 32; do
NWS_TILE_CHUNK=$ch timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --inputs $inp > gpurun_out/b.json 2> gpurun_out/b.err; tail -c 200 gpurun_out/b.err
python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read()); print('$inp chunk $ch', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
done; done
NWS_TILE_CHUNK=8 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --inputs realistic --variant newt > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read()); print('newt realistic', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
