cd /root/repo
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 200 python tests/debug_e2e.py 2>&1 | tail -2
for o in 1 0; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --option cuda_graphs=$o > $O/b1_graph$o.json 2> $O/b1_graph$o.err; python - <<P
import json; d=json.load(open('$O/b1_graph$o.json')); print('graphs $o b1',d['value'],d['e2e']['value'],d['gpu_launches'])
P
tail -2 $O/b1_graph$o.err
done
