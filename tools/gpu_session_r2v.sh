# final 1-GPU verification of round 2: full GPU suite, smoke, default bench + reference arm, ncu launch list, C3 ncu capture
set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_gpu_r2v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2v.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2v.log 2>&1
( time timeout 900 python bench.py > gpurun_out/bench_C3_r2.json 2> gpurun_out/bench_C3_r2.err ) 2> gpurun_out/bench_C3_r2.time
for c in C1 C2 C4; do
timeout 900 python bench.py --config $c --no-all --cpu-seconds 10 > gpurun_out/bench_${c}_r2.json 2> gpurun_out/bench_${c}_r2.err
done
timeout 900 python bench.py --config C5 --no-all --no-e2e --steps 5 --cpu-seconds 10 > gpurun_out/bench_C5_r2.json 2> gpurun_out/bench_C5_r2.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_c3_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-all > gpurun_out/ncu_launches_c3_r2.out 2>&1
timeout 900 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/plain_C3_r2v.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 -f -o gpurun_out/prof_C3_r2v python bench.py --config C3 --steps 2 --warmup 1 --no-cpu --no-e2e --no-all > gpurun_out/ncu_C3_r2v.log 2>&1
timeout 900 python tools/gpu_parity_report.py --n 512 --quad --quad-n 64 --out gpurun_out/parity_r2.json > gpurun_out/parity_r2.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2v.log; cat gpurun_out/smoke_r2v.log | tail -2; cat gpurun_out/bench_C3_r2.time
