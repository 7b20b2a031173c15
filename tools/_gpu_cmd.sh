( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r1k.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r1k.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1k.log 2>&1; echo rc=$? >> gpurun_out/smoke_r1k.log
for c in C1 C2 C3 C4; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_${c}_r1k.json 2> gpurun_out/bench_${c}_r1k.err; done
timeout 300 python bench.py --config C5 --steps 5 --warmup 3 --no-e2e --cpu-seconds 10 > gpurun_out/bench_C5_r1k.json 2> gpurun_out/bench_C5_r1k.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1k.json 2> gpurun_out/bench_ref_r1k.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_r1k.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_r1k.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/plain2_r1k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:inflx_grid -c 1 -o gpurun_out/prof_C3_r1k -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu2_r1k.log 2>&1
