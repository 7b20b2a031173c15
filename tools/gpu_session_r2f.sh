# multi-GPU session: N = number of visible GPUs (gpurun --gpus N)
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r2f_n$N
nvidia-smi topo -m > gpurun_out/topo_$T.txt 2>&1
timeout 300 python tools/numa_probe.py 1 > gpurun_out/numa_probe_$T.txt 2>&1
for numa in 1 0; do
INFLATOX_NUMA=$numa timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$numa bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_C3_numa${numa}_$T.json 2> gpurun_out/bench_C3_numa${numa}_$T.err
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -k "multi_device or long_point or two_streams" -q > gpurun_out/pytest_multidev_$T.log 2>&1
tail -3 gpurun_out/pytest_multidev_$T.log
