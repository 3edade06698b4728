python -m pytest tests/test_gpu_block.py tests/test_gpu_fullsize.py tests/test_gpu_bp.py -x -q 2>&1 | tail -3
echo "== cubic";  python tools/profile_block.py cubic 16 6 5 2>&1 | tail -1
echo "== grid 32 8"; python tools/profile_block.py grid 32 8 20 2>&1 | tail -1
echo "== hh"; python tools/profile_block.py heavyhex 0 32 20 2>&1 | tail -1
