# Launch list of the benchmark command (ncu --metrics gpu__time_duration.sum): profiles/r2_launches_fastnewt.csv
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/b_ncu.log 2>&1
tail -c 200 gpurun_out/b_ncu.log
