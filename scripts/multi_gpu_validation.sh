# 2-GPU validation of the sharded bench and of the distillation step's NCCL gradient all-reduce
cd /root/repo
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_teacher_b1_2gpu.json 2> $O/bench_teacher_b1_2gpu.err; cut -c1-500 $O/bench_teacher_b1_2gpu.json; tail -2 $O/bench_teacher_b1_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload teacher_b16 --steps 5 --warmup 3 > $O/bench_teacher_b16_2gpu.json 2> $O/bench_teacher_b16_2gpu.err; cut -c1-300 $O/bench_teacher_b16_2gpu.json; tail -2 $O/bench_teacher_b16_2gpu.err
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload distill_b1 --steps 10 --warmup 3 > $O/bench_distill_b1_2gpu.json 2> $O/bench_distill_b1_2gpu.err; cut -c1-600 $O/bench_distill_b1_2gpu.json; tail -3 $O/bench_distill_b1_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --impl reference --steps 1 --warmup 3 > $O/bench_reference_2gpu.json 2> $O/bench_reference_2gpu.err; cut -c1-200 $O/bench_reference_2gpu.json
