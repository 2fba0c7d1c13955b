cd /root/repo
O=gpurun_out/final; mkdir -p $O
timeout 200 python tests/debug_e2e.py 2>&1 | tail -2
for m in default tiny; do
if [ $m = default ]; then unset THA4_TC_STAGES; else export THA4_TC_STAGES=$m; fi
timeout 300 python bench.py --workload teacher_b16 --steps 6 --warmup 3 --no-cpu-baseline > $O/b16_$m.json 2> $O/b16_$m.err; python - <<P
import json; d=json.load(open('$O/b16_$m.json')); print('$m b16',d['value'],d['e2e']['value'],d['kernel_time_us_per_step'])
P
done
unset THA4_TC_STAGES
timeout 300 python bench.py --workload teacher_b1_nocache --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_teacher_b1_nocache.json 2> $O/bench_teacher_b1_nocache.err; cut -c1-300 $O/bench_teacher_b1_nocache.json
timeout 400 python bench.py --workload pose_sweep_512 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_pose_sweep_512.json 2> $O/bench_pose_sweep_512.err; cut -c1-300 $O/bench_pose_sweep_512.json; tail -2 $O/bench_pose_sweep_512.err
