nvidia-smi -L | head -10 > gpurun_out/gpus8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_check.py > gpurun_out/multigpu_check_8.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
