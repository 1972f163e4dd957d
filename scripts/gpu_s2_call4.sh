mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"zonal_pair_kernel|zonal_hash_kernel" -s 2 -c 2 -f -o gpurun_out/s2d_zonal python scripts/prof_pair.py 16384 > gpurun_out/s2d_ncu.log 2>&1; tail -3 gpurun_out/s2d_ncu.log
