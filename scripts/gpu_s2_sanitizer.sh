# compute-sanitizer over the kernels rewritten in round 2's second session (small-raster tests)
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 9"
SEL='uniform_kernels or box_path or fused_equals or zonal or majority or crosstab'
timeout 260 $S --tool memcheck python -m pytest tests -q -m gpu -x -k "$SEL and not full_size and not two_gpu and not striped" > gpurun_out/s2_san_memcheck.txt 2>&1; tail -4 gpurun_out/s2_san_memcheck.txt
timeout 260 $S --tool racecheck python -m pytest tests -q -m gpu -x -k "$SEL and not full_size and not two_gpu and not striped" > gpurun_out/s2_san_racecheck.txt 2>&1; tail -4 gpurun_out/s2_san_racecheck.txt
timeout 150 $S --tool synccheck python -m pytest tests -q -m gpu -x -k "uniform_kernels or fused_equals or majority" > gpurun_out/s2_san_synccheck.txt 2>&1; tail -4 gpurun_out/s2_san_synccheck.txt
