python -m pytest tests/test_gpu_prims.py -m gpu -q -x 2>&1 | tail -15
