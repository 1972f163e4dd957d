# historical record of a GPU call of round 2, second session: `scripts/tune/box_sweep2.py` is the earlier version of
# scripts/tune/box_zonal_sweep.py (it could still switch to the first-generation kernel and to 8 / 9 consumer warps).
# second session of round 2, GPU call 1: parity of the new box / pair kernels, A/B sweep, ncu, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/s2_pytest.txt; tail -6 gpurun_out/s2_pytest.txt
timeout 600 python scripts/tune/box_sweep2.py 32768 > gpurun_out/s2_box_sweep.txt 2>&1; tail -40 gpurun_out/s2_box_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:box_stream2 -s 1 -c 2 -f -o gpurun_out/s2_box2 python scripts/prof_conv.py 16384 9,25 > gpurun_out/s2_ncu.log 2>&1; tail -2 gpurun_out/s2_ncu.log
timeout 900 python bench.py --steps 10 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; tail -c 300 gpurun_out/s2_bench.json; tail -3 gpurun_out/s2_bench.err
