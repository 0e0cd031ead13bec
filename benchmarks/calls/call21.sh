set -x
python -m pytest tests/test_gpu_adapter.py tests/test_gpu_head.py tests/test_gpu_decoder.py tests/test_gpu_umma.py -x -q -m gpu 2>&1 | tail -5
python benchmarks/adapter_bench.py > gpurun_out/r2_adapter_bench5.jsonl 2> gpurun_out/r2_adapter_bench5.err; cat gpurun_out/r2_adapter_bench5.jsonl
python benchmarks/one_adapter.py 16 16 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_adapter_launches5.csv python benchmarks/one_adapter.py 16 16 > gpurun_out/r2_adapter_ncu5.log 2>&1
