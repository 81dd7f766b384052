# ncu --set full capture of the NEWT variant of the fused audio kernel (first serial whole-utterance launch), summarised for profiles/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nws_audio_tc_kernel -s 10 -c 1 -f -o gpurun_out/audio_mlp python bench.py --steps 2 --warmup 3 --variant newt --no-cpu-baseline --no-configs > gpurun_out/ncu_audio_mlp.log 2>&1
tail -3 gpurun_out/ncu_audio_mlp.log
python scripts/ncu_summary.py gpurun_out/audio_mlp.ncu-rep gpurun_out/r2_ncu_audio_tc_mlp | head -40
