OUT=gpurun_out/r04g; mkdir -p $OUT
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu > $OUT/t_multi.log 2>&1; echo "multigpu rc=$?"; tail -n 3 $OUT/t_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?"; cut -c1-700 $OUT/bench_n2.json
