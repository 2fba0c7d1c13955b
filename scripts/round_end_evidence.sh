# Round-end evidence run on one B200: smoke, bench lines for every workload, the reference arm, ncu launch list and
# `--set full` captures (summarised into profiles/ afterwards).  The full `pytest -m gpu` suite is run separately.
cd /root/repo
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "error_class" 2>&1 | tail -8 > $O/pytest_error_class.log; cat $O/pytest_error_class.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 600 python bench.py > $O/bench_teacher_b1.json 2> $O/bench_teacher_b1.err; cat $O/bench_teacher_b1.json
for w in teacher_b16 student_b64 distill_b1; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; cut -c1-300 $O/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; cut -c1-300 $O/bench_reference_arm.json
timeout 600 python bench.py --impl torch_cuda --steps 10 --warmup 3 > $O/bench_torch_cuda_eager.json 2> $O/bench_torch_cuda_eager.err; cut -c1-300 $O/bench_torch_cuda_eager.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_teacher_b1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tail_kernel --launch-skip 12 -c 4 -o $O/prof_tail -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_tail.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:conv_tc_kernel|norm_apply_kernel|attention_kernel|conv_igemm_kernel" --launch-skip 1500 -c 24 -o $O/prof_frame_slice -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_slice.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:siren --launch-skip 8 -c 4 -o $O/prof_siren -f python bench.py --workload student_b64 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_siren.log 2>&1
ls -la $O; du -sh gpurun_out
