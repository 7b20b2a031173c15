set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_gpu_r2r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2r.log
timeout 600 python tools/ab.py hyper complete_analysis 16384 '[{"name":"auto(transposed)"},{"name":"direct","store":"direct"},{"name":"direct_late","store":"direct","extra":["-DINFLX_LATE_STORE"]},{"name":"transposed_minb8","minb":8},{"name":"transposed_minb4","minb":4},{"name":"auto_again"}]' 7 > gpurun_out/ab_hyper_r2r.log 2>&1
for m in doc angular egno; do
timeout 600 python tools/ab.py $m complete_analysis 16384 '[{"name":"auto(direct)"},{"name":"transposed","store":"transposed"},{"name":"auto_again"}]' 7 > gpurun_out/ab_${m}_store_r2r.log 2>&1
done
timeout 900 python bench.py > gpurun_out/bench_C3_r2.json 2> gpurun_out/bench_C3_r2.err
timeout 900 python bench.py --config C1 --no-all --cpu-seconds 10 > gpurun_out/bench_C1_r2.json 2> gpurun_out/bench_C1_r2.err
timeout 900 python bench.py --config C5 --no-all --no-e2e --steps 5 --cpu-seconds 10 > gpurun_out/bench_C5_r2.json 2> gpurun_out/bench_C5_r2.err
tail -4 gpurun_out/pytest_gpu_r2r.log
