set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python benchmarks/adapter_bench.py > gpurun_out/r2_adapter_bench7.jsonl 2> gpurun_out/r2_adapter_bench7.err; cat gpurun_out/r2_adapter_bench7.jsonl
python benchmarks/one_adapter.py 16 16 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_adapter_launches7.csv python benchmarks/one_adapter.py 16 16 > gpurun_out/r2_adapter_ncu7.log 2>&1
