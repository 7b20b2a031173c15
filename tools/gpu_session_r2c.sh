set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_random_models.py tests/test_gpu_numerics.py -m gpu -q --timeout=1200 > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c.log
AB='[{"name":"base"},{"name":"atan2","extra":["-DINFLX_ATAN_CHAINS=2"],"check":false},{"name":"atan3","extra":["-DINFLX_ATAN_CHAINS=3"],"check":false},{"name":"early_store","extra":["-DINFLX_EARLY_STORE"]},{"name":"atan2+early","extra":["-DINFLX_ATAN_CHAINS=2","-DINFLX_EARLY_STORE"],"check":false},{"name":"rcp4","extra":["-DINFLX_EXPERIMENT_RCP4"]},{"name":"atan2+early+minb5","minb":5,"extra":["-DINFLX_ATAN_CHAINS=2","-DINFLX_EARLY_STORE"],"check":false},{"name":"base_again"}]'
for m in egno d5 doc; do
timeout 900 python tools/ab.py $m complete_analysis 16384 "$AB" 7 > gpurun_out/ab_${m}_r2c.log 2>&1
done
AB2='[{"name":"base(prepass)"},{"name":"cols_never","cols":"never"},{"name":"cols_never_rpt32","cols":"never","rpt":32},{"name":"prepass_rpt32","rpt":32},{"name":"atan2","extra":["-DINFLX_ATAN_CHAINS=2"],"check":false},{"name":"atan2+early","extra":["-DINFLX_ATAN_CHAINS=2","-DINFLX_EARLY_STORE"],"check":false},{"name":"atan2+early+minb4","minb":4,"extra":["-DINFLX_ATAN_CHAINS=2","-DINFLX_EARLY_STORE"],"check":false},{"name":"atan2+early+minb5","minb":5,"extra":["-DINFLX_ATAN_CHAINS=2","-DINFLX_EARLY_STORE"],"check":false},{"name":"minb4","minb":4},{"name":"libm_cr","libm":"cr","check":false},{"name":"base_again"}]'
timeout 900 python tools/ab.py angular complete_analysis 16384 "$AB2" 7 > gpurun_out/ab_angular_r2c.log 2>&1
timeout 600 python tools/ab.py angular consistency_only 4096 '[{"name":"base(prepass)"},{"name":"cols_never","cols":"never"},{"name":"minb4","minb":4},{"name":"minb6","minb":6},{"name":"base_again"}]' 9 > gpurun_out/ab_angular_con_r2c.log 2>&1
timeout 600 python tools/ab.py hyper complete_analysis 16384 '[{"name":"base"},{"name":"atan2","extra":["-DINFLX_ATAN_CHAINS=2"],"check":false},{"name":"early_store","extra":["-DINFLX_EARLY_STORE"]},{"name":"base_again"}]' 7 > gpurun_out/ab_hyper_r2c.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2c.log
