python -m pytest tests/test_gpu_umma.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
echo "--- TS (default)"; python benchmarks/nce_bench.py 2>&1 | cut -c1-300
echo "--- SS variant"; COR_B200_LIB=cor_b200/build/ab/libcor_b200_ncess.so python benchmarks/nce_bench.py 2>&1 | cut -c1-300
