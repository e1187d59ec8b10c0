(timeout 600 python -m pytest tests/test_eval_ops.py -m gpu -x -q 2>&1 | tail -8)
