set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2e.log
AB='[{"name":"new_default"},{"name":"late_store","extra":["-DINFLX_LATE_STORE"]},{"name":"atan1","extra":["-DINFLX_ATAN_CHAINS=1"],"check":false},{"name":"round1_like","extra":["-DINFLX_ATAN_CHAINS=1","-DINFLX_LATE_STORE"],"check":false},{"name":"rcp4","extra":["-DINFLX_EXPERIMENT_RCP4"]},{"name":"minb6","minb":6},{"name":"minb5","minb":5},{"name":"minb4","minb":4},{"name":"new_default_again"}]'
for m in egno d5 doc angular hyper; do
timeout 900 python tools/ab.py $m complete_analysis 16384 "$AB" 7 > gpurun_out/ab_${m}_r2e.log 2>&1
done
timeout 600 python tools/ab.py angular consistency_only 4096 '[{"name":"new_default"},{"name":"minb5","minb":5},{"name":"minb4","minb":4},{"name":"new_default_again"}]' 9 > gpurun_out/ab_angular_con_r2e.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_C3_r2e.json 2> gpurun_out/bench_C3_r2e.err
tail -5 gpurun_out/pytest_gpu_r2e.log
