set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu_r4.log; cat gpurun_out/pytest_gpu_r4.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b1_r4.json 2> gpurun_out/bench_b1_r4.err; python - <<P
import json; d=json.load(open('gpurun_out/bench_b1_r4.json')); print('b1',d['value'],d['e2e']['value'],d['kernel_time_us_per_step'])
P
timeout 300 python bench.py --workload teacher_b16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b16_r4.json 2> gpurun_out/bench_b16_r4.err; python - <<P
import json; d=json.load(open('gpurun_out/bench_b16_r4.json')); print('b16',d['value'],d['e2e']['value'],d['kernel_time_us_per_step'])
P
timeout 200 python tests/debug_perf4.py 2>&1 | grep "half 1"
