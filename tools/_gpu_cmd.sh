timeout 400 python tools/gpu_parity_report.py --n 512 --cr --models egno d5 angular --out gpurun_out/parity_cr_r1n.json > gpurun_out/parity_cr_r1n.log 2>&1
