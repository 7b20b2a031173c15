# final scaling session of round 2 on ONE 8-GPU box: N = 1, 2, 4, 8 back to back (the driver's protocol)
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python bench.py --no-all --cpu-seconds 5 > gpurun_out/bench_C3_n1_r2t.json 2> gpurun_out/bench_C3_n1_r2t.err
for N in 2 4 8; do
[ $N -le $NG ] || continue
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_C3_n${N}_r2t.json 2> gpurun_out/bench_C3_n${N}_r2t.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --config C5 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_C5_n${NG}_r2t.json 2> gpurun_out/bench_C5_n${NG}_r2t.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -k "multi_device or long_point or two_streams" -q > gpurun_out/pytest_multidev_r2t.log 2>&1
tail -3 gpurun_out/pytest_multidev_r2t.log
for N in 1 2 4 8; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_C3_n${N}_r2t.json"))
    print($N, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("e2e_inprocess") or {}).get("value"))
except Exception as e: print($N, "ERR", e)
PY
done
