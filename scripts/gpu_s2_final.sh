# round 2, second session: final single-GPU evidence (tests, bench line, reference arm, ncu launch list and captures)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/s2f_pytest.txt; tail -6 gpurun_out/s2f_pytest.txt
timeout 900 python bench.py --steps 10 > gpurun_out/s2f_bench.json 2> gpurun_out/s2f_bench.err; tail -c 400 gpurun_out/s2f_bench.json; tail -3 gpurun_out/s2f_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s2f_bench_ref.json 2> gpurun_out/s2f_bench_ref.err; tail -c 300 gpurun_out/s2f_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2f_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-host --skip-ops > gpurun_out/s2f_launches.log 2>&1; tail -2 gpurun_out/s2f_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:box_stream2 -s 1 -c 3 -f -o gpurun_out/s2f_box2 python scripts/prof_conv.py 16384 9,25 > gpurun_out/s2f_ncu_box.log 2>&1; tail -2 gpurun_out/s2f_ncu_box.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"zonal_pair_kernel|zonal_hash_kernel" -s 2 -c 2 -f -o gpurun_out/s2f_zonal python scripts/prof_pair.py 16384 > gpurun_out/s2f_ncu_zonal.log 2>&1; tail -2 gpurun_out/s2f_ncu_zonal.log
