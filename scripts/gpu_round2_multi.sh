# 2-GPU check: the GPU test-suite (the multi-GPU tests run), then the bench at N = 2 under torchrun
N=${1:-2}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_n$N.txt; tail -6 gpurun_out/pytest_gpu_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
