# Round evidence run on one B200 (gpurun): tests, bench lines, launch list, ncu captures, parity report.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 300 gpurun_out/bench_fast.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fast_k20.json 2> gpurun_out/bench_fast_k20.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/b_ncu.log 2>&1
# serial forwards (scripts/dev_serial_forward.py): one whole-batch launch per kernel; the 3rd forward's launches are captured
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nws_audio_tc_kernel -s 2 -c 1 -f -o gpurun_out/audio_lut python scripts/dev_serial_forward.py fastnewt 4 > gpurun_out/ncu_audio.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nws_audio_tc_kernel -s 2 -c 1 -f -o gpurun_out/audio_mlp python scripts/dev_serial_forward.py newt 4 > gpurun_out/ncu_audio_mlp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"nws_mlp_tc_kernel|nws_noise_filter_kernel|nws_reverb_|nws_gru_mma_kernel|nws_phase_carry|nws_noise_spectrum|nws_rng" -s 16 -c 8 -f -o gpurun_out/r2_hop_kernels python scripts/dev_serial_forward.py fastnewt 4 > gpurun_out/ncu_hop.log 2>&1
timeout 300 python scripts/parity_report.py --json gpurun_out/parity.json 2>&1 | tail -3
PYTHONPATH=. timeout 300 python scripts/dev_sweep.py > gpurun_out/sweep.log 2>&1; tail -2 gpurun_out/sweep.log
PYTHONPATH=. timeout 300 python scripts/time_streaming.py --json gpurun_out/streaming.json > gpurun_out/streaming.log 2>&1; tail -2 gpurun_out/streaming.log
PYTHONPATH=. timeout 300 python scripts/time_loudness.py --json gpurun_out/loudness.json 2>&1 | tail -1
python - <<'P'
import json
for f in ("bench_fast", "bench_fast_k20", "bench_ref"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read())
        print(f, d.get("ms_per_step"), d.get("value"), d.get("latency"), d.get("e2e"), d.get("clocks"), (d.get("roofline") or {}).get("kernel_ms"), d.get("stages_ms"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "FAILED", e)
P
