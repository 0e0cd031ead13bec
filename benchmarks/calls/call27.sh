python -m pytest tests/test_gpu_head.py tests/test_gpu_adapter.py -x -q -m gpu 2>&1 | tail -3
python benchmarks/adapter_bench.py --cases 16x16,16x64 > gpurun_out/r2_adapter_bench9.jsonl 2> gpurun_out/r2_adapter_bench9.err; cat gpurun_out/r2_adapter_bench9.jsonl
python benchmarks/one_adapter.py 16 16 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_adapter_launches9.csv python benchmarks/one_adapter.py 16 16 > gpurun_out/r2_adapter_ncu9.log 2>&1
