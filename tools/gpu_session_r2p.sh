# closing session of round 2: tests, parity report, bench lines, ncu launch list + full captures
set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_gpu_r2p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2p.log
timeout 900 python tools/gpu_parity_report.py --n 512 --quad --quad-n 64 --out gpurun_out/parity_r2.json > gpurun_out/parity_r2.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2p.log 2>&1
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
timeout 900 python bench.py > gpurun_out/bench_C3_r2.json 2> gpurun_out/bench_C3_r2.err
for c in C2 C4; do
timeout 900 python bench.py --config $c --no-all --cpu-seconds 10 > gpurun_out/bench_${c}_r2.json 2> gpurun_out/bench_${c}_r2.err
done
timeout 300 python tools/scalar_latency.py --n 10000 2>/dev/null | grep calc_V > gpurun_out/scalar_r2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_c3_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-all > gpurun_out/ncu_launches_c3_r2.out 2>&1
for c in C3 C4 C6; do
timeout 900 python bench.py --config $c --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/plain_${c}_r2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 -f -o gpurun_out/prof_${c}_r2 python bench.py --config $c --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/ncu_${c}_r2.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -5
tail -3 gpurun_out/pytest_gpu_r2p.log
