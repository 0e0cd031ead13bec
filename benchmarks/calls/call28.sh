for r in 64 128; do echo "rows=$r"; COR_SEG_STRIP_ROWS=$r python benchmarks/seg_bench.py --batches 64,128 2>&1 | grep '"kernel": "strip"' | cut -c1-200; done
COR_SEG_STRIP_ROWS=128 python -m pytest tests/test_gpu_seg.py -q -m gpu 2>&1 | tail -2
