set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r2h_n$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_C3_$T.json 2> gpurun_out/bench_C3_$T.err
tail -c 600 gpurun_out/bench_C3_$T.json
