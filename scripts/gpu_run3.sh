set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_distill.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_distill_face.log; cat gpurun_out/pytest_distill_face.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_teacher_b16_v5.csv python bench.py --workload teacher_b16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b16.log 2>&1
tail -2 gpurun_out/ncu_b16.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_teacher_b1_v5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
tail -2 gpurun_out/ncu_b1.log | cut -c1-300
