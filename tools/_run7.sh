python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -2 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tests/dist_gpu_check.py > gpurun_out/dist_check_2gpu.log 2>&1; tail -3 gpurun_out/dist_check_2gpu.log
python tests/dist_gpu_check.py --threads 2 > gpurun_out/dist_check_threads2.log 2>&1; tail -2 gpurun_out/dist_check_threads2.log
