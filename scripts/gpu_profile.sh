#!/bin/bash
# the profile artefacts of a round (keep gpurun_out under 64 MiB: it is not copied back otherwise)
mkdir -p gpurun_out
if [ "${LIST:-1}" = 1 ]; then
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline --no-fp32-variant --layers-out gpurun_out/layers_ncu.json > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_tf32.csv
fi
# conv_tc2 launches in program order: 0 stem | s0b0 1-4 (1x1a, 3x3, 1x1b, proj+res) | s1b1 15-17 | s2b1 28-30 | s3b1 47-49 ; the first 54 belong to the cold first call
for spec in ${SPECS-0:1 2:1 4:1}; do   # SPECS="" skips the full captures (3 reports fit the 64 MiB gpurun_out limit)
  s=${spec%%:*}; c=${spec##*:}
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:conv_tc2|conv_patch' -s $((54 + s)) -c $c -o gpurun_out/prof_tc2_l$s -f python bench.py --steps 1 --warmup 3 --precision ${PREC:-tf32} --no-cpu-baseline --no-fp32-variant > gpurun_out/ncu_full_l$s.log 2>&1
done
# the reports are ~20 MB each: keep their metrics (markdown) and source pages (gzip-ed csv), drop the reports themselves
if [ -n "${MD:-}" ]; then
  args=(); for spec in ${SPECS-0:1 2:1 4:1}; do s=${spec%%:*}; args+=("${PREC_TITLE:-TF32} conv launch $s" gpurun_out/prof_tc2_l$s.ncu-rep); done
  python scripts/ncu_full_md.py gpurun_out/$MD "${args[@]}"; rm -f gpurun_out/*.ncu-rep
fi
ls -la gpurun_out | tail -20; du -sh gpurun_out
