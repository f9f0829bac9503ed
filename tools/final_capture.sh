#!/bin/bash
# One GPU call that refreshes every measured artefact of a round (run from the repo root through gpurun, ~8 GPU-minutes):
#   gpurun --timeout 1500 -- 'bash tools/final_capture.sh r02'
# then, back in the build container:
#   cp gpurun_out/$TAG/${TAG}_* profiles/   (the .ncu-rep files stay in gpurun_out; their summaries are what gets committed)
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$2" != "nobench" ]; then
python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; cut -c1-200 $OUT/${TAG}_bench_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
fi
# ncu launch list of ONE eager training step (cold-cache, serialised: shares of the step, not absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches_eager_step.csv python bench.py --profile-step --warmup 3 > $OUT/ncu_step.log 2>&1
python tools/agg_launches.py $OUT/${TAG}_launches_eager_step.csv 60 > $OUT/${TAG}_launches_eager_step_summary.txt
# full metric set: one launch of every kernel class of the real step (first occurrence of each name inside one eager step)
ncu --set full --clock-control none --profile-from-start off \
    -k regex:"gemm_tc_kernel|region_aggregate|norm_fwd_vec|norm_bwd_vec|cast_f32_bf16|lstm_cell|lstm_step|attn2|adam_multi|ce_masked|latent_psl" \
    --kernel-id :::1 -o $OUT/${TAG}_step_kernels python bench.py --profile-step --warmup 3 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/${TAG}_step_kernels.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null && \
    python tools/ncu_summary.py $OUT/raw.csv > $OUT/${TAG}_ncu_full_step_kernels.json 2> $OUT/ncu_summary.err
# the dominant GEMMs / streaming kernels in isolation at the benched shapes
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|norm_bwd_bf16|norm_fwd_bf16|cast_f32_bf16" \
    -o $OUT/${TAG}_top_kernels python tools/profile_kernels.py > $OUT/ncu_full2.log 2>&1
ncu -i $OUT/${TAG}_top_kernels.ncu-rep --page raw --csv > $OUT/raw2.csv 2>/dev/null && \
    python tools/ncu_summary.py $OUT/raw2.csv > $OUT/${TAG}_ncu_full_top_kernels.json
# warm per-kernel roofline fractions, fused aggregation kernels, vocabulary-row kernels, per-block step breakdown, GAN iteration
python tools/roofline_table.py > $OUT/${TAG}_roofline_table.jsonl 2> $OUT/roofline.err
python tools/bench_region_agg.py > $OUT/${TAG}_region_agg.jsonl 2> $OUT/region_agg.err
python tools/bench_vocab_rows.py > $OUT/${TAG}_vocab_rows.jsonl 2> $OUT/vocab_rows.err
python tools/profile_blocks.py > $OUT/${TAG}_blocks.jsonl 2> $OUT/blocks.err
python tools/bench_gan.py > $OUT/${TAG}_gan_iteration.json 2> $OUT/gan.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/dstep.csv python tools/bench_gan.py --profile-dstep > $OUT/ncu_dstep.log 2>&1
python tools/agg_launches.py $OUT/dstep.csv 60 > $OUT/${TAG}_launches_dstep_summary.txt 2>/dev/null
tail -5 $OUT/gan.err $OUT/ncu_dstep.log
# the reports themselves are too large to travel back (gpurun_out is capped at 64 MiB): keep the summaries
gzip -f $OUT/raw.csv; rm -f $OUT/*.ncu-rep $OUT/raw2.csv
ls -la $OUT
