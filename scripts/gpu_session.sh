timeout 1500 python -m pytest tests/test_gpu_statistical.py tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -12
