set -x
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_pdl.log; cat gpurun_out/pytest_gpu_pdl.log
for o in 1 0; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --option pdl=$o > gpurun_out/bench_b1_pdl$o.json 2> gpurun_out/bench_b1_pdl$o.err; python - <<P
import json; d=json.load(open('gpurun_out/bench_b1_pdl$o.json')); print('pdl',$o,'b1',d['value'],d['e2e']['value'],d['kernel_time_us_per_step'])
P
done
timeout 300 python bench.py --workload teacher_b16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b16_pdl.json 2> gpurun_out/bench_b16_pdl.err; python - <<P
import json; d=json.load(open('gpurun_out/bench_b16_pdl.json')); print('b16',d['value'],d['e2e']['value'],d['kernel_time_us_per_step'])
P
tail -3 gpurun_out/*pdl*.err
