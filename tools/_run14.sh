python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_gpu_tests.log; cat gpurun_out/final_gpu_tests.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
