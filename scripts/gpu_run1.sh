set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" 2>&1 | tail -15 > gpurun_out/pytest_f16_conv.log
cat gpurun_out/pytest_f16_conv.log
timeout 300 python tests/debug_perf4.py > gpurun_out/perf4.log 2>&1; cat gpurun_out/perf4.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_f16.log; cat gpurun_out/pytest_gpu_f16.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b1_f16.json 2> gpurun_out/bench_b1_f16.err; cat gpurun_out/bench_b1_f16.json
timeout 300 python bench.py --workload teacher_b16 --steps 5 --warmup 3 > gpurun_out/bench_b16_f16.json 2> gpurun_out/bench_b16_f16.err; cat gpurun_out/bench_b16_f16.json
