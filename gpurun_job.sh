python -m pytest tests/test_gpu_rule_shims.py -m gpu -x -q 2>&1 | tail -15
