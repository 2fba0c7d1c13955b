cd /root/repo
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke_final.log
timeout 600 python bench.py > $O/bench_teacher_b1_final.json 2> $O/bench_teacher_b1_final.err; python - <<P
import json; d=json.load(open('$O/bench_teacher_b1_final.json')); print('b1',d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline_tail']['frac'],d['cpu_baseline']['value'])
P
timeout 300 python bench.py --workload pose_sweep_512 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_pose_sweep_512_final.json 2> $O/bench_pose_sweep_512_final.err; cut -c1-160 $O/bench_pose_sweep_512_final.json
