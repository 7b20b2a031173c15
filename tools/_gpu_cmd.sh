( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r1l.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r1l.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default_r1l.json 2> gpurun_out/bench_default_r1l.err
for c in C1 C2; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_${c}_r1l.json 2> gpurun_out/bench_${c}_r1l.err; done
for c in C1 C2 C4; do
  timeout 300 python bench.py --config $c --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/plain_${c}_r1l.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 -o gpurun_out/prof_${c}_r1l -f python bench.py --config $c --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_${c}_r1l.log 2>&1
done
timeout 400 python tools/gpu_parity_report.py --n 512 --quad --quad-n 64 --out gpurun_out/parity_r1l.json > gpurun_out/parity_r1l.log 2>&1
