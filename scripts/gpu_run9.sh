cd /root/repo
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_teacher_b1_v2.json 2> $O/bench_teacher_b1_v2.err; python - <<P
import json; d=json.load(open('$O/bench_teacher_b1_v2.json')); print('b1',d['value'],d['e2e']['value'],d['kernel_time_us_per_step'], d['roofline_tail'])
P
timeout 600 python bench.py --no-cpu-baseline --workload teacher_b16 --steps 10 --warmup 3 > $O/bench_teacher_b16_v2.json 2> $O/bench_teacher_b16_v2.err; python - <<P
import json; d=json.load(open('$O/bench_teacher_b16_v2.json')); print('b16',d['value'],d['e2e']['value'],d['kernel_time_us_per_step'], d['roofline_tail'])
P
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tail_kernel --launch-skip 12 -c 4 -o $O/prof_tail_v2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_tail_v2.log 2>&1
du -sh gpurun_out
