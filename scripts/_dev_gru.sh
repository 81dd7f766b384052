mkdir -p gpurun_out
# serial forward = launches: rng, phase_carry, noise_spectrum, gru, mlp_tc, noise_filter, audio, reverb x3 (+ first-call extras)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nws_mlp_tc_kernel|nws_noise_filter_kernel|nws_reverb_" -s 10 -c 5 -f -o gpurun_out/r2_hop_kernels python scripts/dev_serial_forward.py fastnewt 4 > gpurun_out/ncu_hop.log 2>&1
tail -3 gpurun_out/ncu_hop.log
