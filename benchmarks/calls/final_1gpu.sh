# Round-end evidence on one B200: every line lands in gpurun_out/r02_final_* (copied into profiles/ afterwards).
O=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/r02_final_pytest.log; cat $O/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_final_smoke.log 2>&1; tail -2 $O/r02_final_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_final_reference_arm.json 2> $O/r02_final_reference_arm.err
python bench.py > $O/r02_final_bench_n1.json 2> $O/r02_final_bench_n1.err; cut -c1-600 $O/r02_final_bench_n1.json
python benchmarks/sweep.py > $O/r02_final_sweep.jsonl 2> $O/r02_final_sweep.err
python benchmarks/seg_bench.py > $O/r02_final_seg_bench.jsonl 2> $O/r02_final_seg_bench.err
python benchmarks/nce_bench.py > $O/r02_final_nce_bench.jsonl 2> $O/r02_final_nce_bench.err
python benchmarks/sim_bench.py > $O/r02_final_sim_bench.jsonl 2> $O/r02_final_sim_bench.err
python benchmarks/adapter_bench.py > $O/r02_final_adapter_bench.jsonl 2> $O/r02_final_adapter_bench.err
ls -la $O | grep r02_final
