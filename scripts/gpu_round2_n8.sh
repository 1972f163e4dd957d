# bench at N GPUs under torchrun only (no pytest): the driver's scaling step, run by hand once
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; grep "^{" gpurun_out/bench_n$N.json | tail -c 2500; tail -5 gpurun_out/bench_n$N.err
