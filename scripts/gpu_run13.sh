cd /root/repo
O=gpurun_out/final; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
for mb in 16 32; do
timeout 400 python bench.py --workload pose_sweep_512 --steps 2 --warmup 3 --no-cpu-baseline --option microbatch=$mb > $O/sweep_mb$mb.json 2> $O/sweep_mb$mb.err; python - <<P
import json
try:
    d=json.load(open('$O/sweep_mb$mb.json')); print('microbatch $mb sweep512',d['value'],d['e2e']['value'])
except Exception as e: print('mb $mb failed', e)
P
tail -2 $O/sweep_mb$mb.err
done
timeout 300 python bench.py --workload teacher_b16 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_teacher_b16_final.json 2> $O/bench_teacher_b16_final.err; cut -c1-200 $O/bench_teacher_b16_final.json
