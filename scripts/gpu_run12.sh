cd /root/repo
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 200 python tests/debug_e2e.py 2>&1 | tail -1
for mb in 4 16; do
timeout 300 python bench.py --workload teacher_b16 --steps 6 --warmup 3 --no-cpu-baseline --option microbatch=$mb > $O/b16_mb$mb.json 2> $O/b16_mb$mb.err; python - <<P
import json
try:
    d=json.load(open('$O/b16_mb$mb.json')); print('microbatch $mb b16',d['value'],d['e2e']['value'])
except Exception as e: print('mb $mb failed', e)
P
tail -2 $O/b16_mb$mb.err
done
timeout 600 python bench.py > $O/bench_teacher_b1_final.json 2> $O/bench_teacher_b1_final.err; cat $O/bench_teacher_b1_final.json | cut -c1-700
