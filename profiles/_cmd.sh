timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k pair 2>&1 | tail -12
