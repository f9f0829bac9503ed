OUT=gpurun_out/r04f; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 $OUT/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_fused.json 2> $OUT/bench_fused.err; echo "bench rc=$?"
DLSG_FUSED_CELL_ATTN=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_unfused.json 2> $OUT/bench_unfused.err; echo "bench rc=$?"
python - <<'PY'
import json
for k in ('fused','unfused'):
    try:
        d=json.load(open('gpurun_out/r04f/bench_%s.json'%k))
        print(k, 'step %.3f e2e %.3f tf0.8 %.3f greedy %.0f beam %.0f gan %.2f launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['graphed_scheduled_sampling_tf0.8_ms_per_step'], d['greedy_captions_per_s_B256'], d['beam5_captions_per_s_B128'], d['gan_iteration_ms_B64'], d['gpu_launches']))
    except Exception as e: print(k, 'ERR', e)
PY
