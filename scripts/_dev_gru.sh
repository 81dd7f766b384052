mkdir -p gpurun_out
for cfg in "32 117" "32 156" "32 234" "64 218" "16 121" "125 125" "250 250" "32 94"; do set -- $cfg; NWS_PIPE_FIRST=$1 NWS_PIPE_BLOCK=$2 timeout 200 python scripts/dev_lanes.py fastnewt 64 2>&1 | grep -v Warn | tee -a gpurun_out/dev_lanes.log; done
for ch in 4 16; do NWS_TILE_CHUNK=$ch timeout 200 python scripts/dev_lanes.py fastnewt 64 2>&1 | grep -v Warn | tee -a gpurun_out/dev_lanes.log; done
for cfg in "32 117" "32 234" "125 125"; do set -- $cfg; NWS_PIPE_FIRST=$1 NWS_PIPE_BLOCK=$2 timeout 200 python scripts/dev_lanes.py newt 64 2>&1 | grep -v Warn | tee -a gpurun_out/dev_lanes.log; done
