python -m pytest tests/test_gpu_match.py tests/test_gpu_diag.py -m gpu -x -q 2>&1 | tail -8
