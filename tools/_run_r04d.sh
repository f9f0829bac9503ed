OUT=gpurun_out/r04d; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "lstm" > $OUT/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_graphs_gpu.py -x -q -m gpu -k "disc or gan" > $OUT/t_disc.log 2>&1; echo "disc rc=$?"
tail -n 3 $OUT/t_kernels.log $OUT/t_disc.log
timeout 300 python tools/bench_gan.py > $OUT/gan_iteration.json 2> $OUT/gan.err; cat $OUT/gan_iteration.json
timeout 300 python tools/bench_gan.py --overlap-g 0 > $OUT/gan_iteration_nooverlap.json 2> $OUT/gan0.err; cat $OUT/gan_iteration_nooverlap.json
