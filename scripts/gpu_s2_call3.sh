# historical record of a GPU call of round 2, second session: `scripts/tune/box_sweep2.py` is the earlier version of
# scripts/tune/box_zonal_sweep.py (it could still switch to the first-generation kernel and to 8 / 9 consumer warps).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/s2c_pytest.txt; tail -6 gpurun_out/s2c_pytest.txt
timeout 300 python scripts/bench_zonal.py > gpurun_out/s2c_bench_zonal.txt 2>&1; tail -8 gpurun_out/s2c_bench_zonal.txt
XRS_SWEEP_SHORT=1 timeout 900 python scripts/tune/box_sweep2.py 32768 > gpurun_out/s2c_box_sweep.txt 2>&1; tail -30 gpurun_out/s2c_box_sweep.txt
