( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r1o.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r1o.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1o.log 2>&1; echo rc=$? >> gpurun_out/smoke_r1o.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default_r1o.json 2> gpurun_out/bench_default_r1o.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_r1o.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1o.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_r1o.log 2>&1
