mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest.log
timeout 300 python scripts/dev_pipe.py 2>&1 | grep -v Warning | tee gpurun_out/dev_pipe.log
