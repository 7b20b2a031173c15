set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_gpu_r2q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2q.log
timeout 600 python tools/gpu_parity_report.py --n 512 --models hyper --out gpurun_out/parity_hyper_r2q.json > gpurun_out/parity_hyper_r2q.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_C3_r2.json 2> gpurun_out/bench_C3_r2.err
timeout 900 python bench.py --config C1 --no-all --cpu-seconds 10 > gpurun_out/bench_C1_r2.json 2> gpurun_out/bench_C1_r2.err
timeout 900 python bench.py --config C5 --no-all --no-e2e --steps 5 --cpu-seconds 10 > gpurun_out/bench_C5_r2.json 2> gpurun_out/bench_C5_r2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_c3_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-all > gpurun_out/ncu_launches_c3_r2.out 2>&1
timeout 900 python bench.py --config C7 --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/plain_C7_r2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 -f -o gpurun_out/prof_C7_r2 python bench.py --config C7 --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/ncu_C7_r2.log 2>&1
timeout 600 python tools/ab.py hyper complete_analysis 16384 '[{"name":"default"},{"name":"late_store","extra":["-DINFLX_LATE_STORE"]},{"name":"minb8","minb":8},{"name":"default_again"}]' 7 > gpurun_out/ab_hyper_r2q.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2q.log
