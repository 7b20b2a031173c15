set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_numerics.py tests/test_gpu_random_models.py tests/test_gpu_parity.py -m gpu -q --timeout=1500 > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2g.log
AB='[{"name":"default"},{"name":"chk_fmnmx","extra":["-DINFLX_CHK_FMNMX"]},{"name":"late_store","extra":["-DINFLX_LATE_STORE"]},{"name":"atan2_estrin","extra":["-DINFLX_ATAN_CHAINS=2"],"check":false},{"name":"rcp4","extra":["-DINFLX_EXPERIMENT_RCP4"]},{"name":"default_again"}]'
for m in egno d5 doc angular hyper; do
timeout 900 python tools/ab.py $m complete_analysis 16384 "$AB" 7 > gpurun_out/ab_${m}_r2g.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_C3_r2g.json 2> gpurun_out/bench_C3_r2g.err
tail -3 gpurun_out/pytest_gpu_r2g.log
