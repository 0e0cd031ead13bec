# usage: final_ngpu.sh N   -- the driver's scaling line at N GPUs + the config-4 (100 masks) line + the peer-exchange tests
N=$1; O=gpurun_out
python -m pytest tests/test_gpu_peer.py -q -m gpu 2>&1 | tail -3 > $O/r02_final_peer_tests_n$N.log; cat $O/r02_final_peer_tests_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_final_bench_n$N.json 2> $O/r02_final_bench_n$N.err; cut -c1-400 $O/r02_final_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --masks 100 --no-e2e --no-cpu-baseline --no-u8-variant > $O/r02_final_bench_n${N}_m100.json 2> $O/r02_final_bench_n${N}_m100.err; cut -c1-400 $O/r02_final_bench_n${N}_m100.json
