mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nws_gru_mma_kernel -s 2 -c 1 -f -o gpurun_out/gru_mma python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_gru_mma.log 2>&1
tail -3 gpurun_out/ncu_gru_mma.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gru or pipelined or full_size" 2>&1 | tail -3
