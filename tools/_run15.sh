export ITN_BLOCK_SERIAL=1
ncu --set full --clock-control none --import-source on -k regex:k_block -s 32 -c 3 -f -o gpurun_out/kb_cubic_final2 python tools/profile_block.py cubic 8 6 2 2>&1 | tail -1
unset ITN_BLOCK_SERIAL
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gate_launches_final.csv python tools/profile_gates.py 64 16 4 2>&1 | tail -2
