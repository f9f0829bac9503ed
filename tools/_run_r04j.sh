OUT=gpurun_out/r04j; mkdir -p $OUT
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_graphs_gpu.py -x -q -m gpu -k "disc or gan" > $OUT/t_disc.log 2>&1; echo "disc rc=$?"; tail -n 3 $OUT/t_disc.log
timeout 300 python tools/bench_gan.py > $OUT/gan_iteration.json 2> $OUT/gan.err; cat $OUT/gan_iteration.json
DLSG_SMALL_BMM_SIMT=0 timeout 300 python tools/bench_gan.py > $OUT/gan_iteration_tc.json 2> $OUT/gan_tc.err; cat $OUT/gan_iteration_tc.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/dstep.csv python tools/bench_gan.py --profile-dstep > $OUT/ncu_dstep.log 2>&1
python tools/agg_launches.py $OUT/dstep.csv 60 > $OUT/launches_dstep_summary.txt 2>/dev/null; head -14 $OUT/launches_dstep_summary.txt
gzip -f $OUT/dstep.csv
