mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/s2g_pytest.txt; tail -6 gpurun_out/s2g_pytest.txt
timeout 900 python bench.py --steps 10 > gpurun_out/s2g_bench.json 2> gpurun_out/s2g_bench.err; tail -c 300 gpurun_out/s2g_bench.json; tail -3 gpurun_out/s2g_bench.err
