python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_gpu_tests.log; cat gpurun_out/final_gpu_tests.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
export ITN_BLOCK_SERIAL=1
ncu --set full --clock-control none --import-source on -k regex:k_block -s 32 -c 3 -f -o gpurun_out/kb_cubic_final python tools/profile_block.py cubic 8 6 2 2>&1 | tail -1
