set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi_r2a.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
timeout 600 python tools/gpu_parity_report.py --n 512 --out gpurun_out/parity_r2a.json > gpurun_out/parity_r2a.log 2>&1
timeout 300 python tools/scalar_latency.py --n 5000 > gpurun_out/scalar_r2a.log 2>&1
for m in egno angular d5 hyper; do
timeout 600 python tools/tune.py $m complete_analysis 16384 '[{"rpt":16,"block":128,"minb":6},{"rpt":16,"block":128,"minb":5},{"rpt":16,"block":128,"minb":4},{"rpt":16,"block":128,"minb":6,"extra":["-DINFLX_EXPERIMENT_RCP4"]},{"rpt":16,"block":128,"minb":6,"extra":["-DINFLX_EXPERIMENT_ATAN_VOTE"]},{"rpt":16,"block":128,"minb":6,"extra":["-DINFLX_EXPERIMENT_RCP4","-DINFLX_EXPERIMENT_ATAN_VOTE"]},{"rpt":16,"block":128,"minb":6,"libm":"glibc-all"},{"rpt":16,"block":128,"minb":6,"libm":"cr"}]' > gpurun_out/tune_${m}_r2a.log 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C3_r2a.json 2> gpurun_out/bench_C3_r2a.err
tail -3 gpurun_out/pytest_gpu_r2a.log
