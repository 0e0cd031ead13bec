python -m pytest tests/test_gpu_adapter.py -x -q -m gpu -k depthwise 2>&1 | tail -40
