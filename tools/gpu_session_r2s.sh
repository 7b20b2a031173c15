set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_gpu_r2s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2s.log
timeout 600 python tools/rcp4_device_check.py 42 > gpurun_out/rcp4_check_r2s.txt 2>&1
for m in egno d5 doc angular; do
timeout 600 python tools/ab.py $m complete_analysis 16384 '[{"name":"default(4-FMA reciprocal)"},{"name":"nvcc 5-FMA reciprocal","extra":["-DINFLX_RCP_NVCC"]},{"name":"default_again"}]' 7 > gpurun_out/ab_${m}_r2s.log 2>&1
done
timeout 900 python tools/gpu_parity_report.py --n 512 --quad --quad-n 64 --out gpurun_out/parity_r2.json > gpurun_out/parity_r2.log 2>&1
( time timeout 900 python bench.py > gpurun_out/bench_C3_r2.json 2> gpurun_out/bench_C3_r2.err ) 2> gpurun_out/bench_C3_r2.time
for c in C1 C2 C4; do
timeout 900 python bench.py --config $c --no-all --cpu-seconds 10 > gpurun_out/bench_${c}_r2.json 2> gpurun_out/bench_${c}_r2.err
done
timeout 900 python bench.py --config C5 --no-all --no-e2e --steps 5 --cpu-seconds 10 > gpurun_out/bench_C5_r2.json 2> gpurun_out/bench_C5_r2.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_c3_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-all > gpurun_out/ncu_launches_c3_r2.out 2>&1
for c in C3 C4 C6 C7; do
timeout 900 python bench.py --config $c --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/plain_${c}_r2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 -f -o gpurun_out/prof_${c}_r2s python bench.py --config $c --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/ncu_${c}_r2.log 2>&1
done
tail -4 gpurun_out/pytest_gpu_r2s.log
