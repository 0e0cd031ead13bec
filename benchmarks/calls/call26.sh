set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python benchmarks/adapter_bench.py > gpurun_out/r2_adapter_bench8.jsonl 2> gpurun_out/r2_adapter_bench8.err; cat gpurun_out/r2_adapter_bench8.jsonl
python benchmarks/one_adapter.py 16 16 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_adapter_launches8.csv python benchmarks/one_adapter.py 16 16 > gpurun_out/r2_adapter_ncu8.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_umma_kernel -s 3 -c 2 -o gpurun_out/r02_prof_adapter_gemm2 -f python benchmarks/one_adapter.py 16 16 > gpurun_out/r2_adapter_ncu8b.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"act_bwd_vec_kernel<__nv_bfloat16|dwconv_cl_kernel<false>" -c 2 -o gpurun_out/r02_prof_adapter_ew2 -f python benchmarks/one_adapter.py 16 16 > gpurun_out/r2_adapter_ncu8c.log 2>&1
