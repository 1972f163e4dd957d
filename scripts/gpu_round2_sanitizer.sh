# compute-sanitizer over the GPU parity suite (small-raster tests only), round 2
S="compute-sanitizer --error-exitcode 9"
SEL='not full_size and not two_gpu and not chunked_rows and not several_devices and not many_tiles and not pinned_cache'
timeout 1500 $S --tool memcheck python -m pytest tests -q -m gpu -x -k "$SEL" > gpurun_out/r02_san_memcheck.txt 2>&1; tail -4 gpurun_out/r02_san_memcheck.txt
timeout 1500 $S --tool racecheck python -m pytest tests -q -m gpu -x -k "reference_outputs or tma_path_vs_oracle or uniform_kernels or box_path or zonal or majority or crosstab or geodesic or ingest or focal_stats_tma" > gpurun_out/r02_san_racecheck.txt 2>&1; tail -4 gpurun_out/r02_san_racecheck.txt
timeout 900 $S --tool synccheck python -m pytest tests -q -m gpu -x -k "reference_outputs or uniform_kernels or zonal or majority or crosstab or geodesic" > gpurun_out/r02_san_synccheck.txt 2>&1; tail -4 gpurun_out/r02_san_synccheck.txt
