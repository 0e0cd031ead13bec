for r in 64 128; do echo "rows=$r"; COR_SEG_STRIP_ROWS=$r python benchmarks/seg_bench.py --batches 64,128 2>&1 | grep '"kernel": "strip"' | cut -c1-200; done
COR_SEG_STRIP_ROWS=128 python -m pytest tests/test_gpu_seg.py -q -m gpu 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python benchmarks/adapter_bench.py > gpurun_out/r2_adapter_bench10.jsonl 2> gpurun_out/r2_adapter_bench10.err; cat gpurun_out/r2_adapter_bench10.jsonl
python benchmarks/one_adapter.py 16 16 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_adapter_launches10.csv python benchmarks/one_adapter.py 16 16 > gpurun_out/r2_adapter_ncu10.log 2>&1
