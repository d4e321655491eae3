#!/bin/bash
# round-2 ncu captures (one launch each, --set full) of the kernels that were not re-profiled with the bench: run under gpurun
set -x
python scripts/quick_bench.py 16384 C-SPEC:tf,R-SPEC:tf,R-FBANK,R-MFCC > gpurun_out/r2q_plain.log 2>&1 || exit 1
for spec in "C-SPEC:tf cspec_tf" "R-SPEC:tf r_spec_tf" "R-FBANK r_fbank" "R-MFCC r_mfcc"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:srfe_ -s 3 -c 1 -o gpurun_out/prof_r2_$2 python scripts/quick_bench.py 16384 $1 > gpurun_out/r2q_ncu_$2.log 2>&1
done
python scripts/fused_once.py > gpurun_out/r2q_fused_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:srfe_spec_fbank -s 2 -c 1 -o gpurun_out/prof_r2_fused_spec_fbank python scripts/fused_once.py > gpurun_out/r2q_ncu_fused.log 2>&1
for n in cspec_tf r_spec_tf r_fbank r_mfcc fused_spec_fbank; do
  bash scripts/profile_summary.sh gpurun_out/prof_r2_$n.ncu-rep gpurun_out/r2_ncu_${n}_16384.txt
  rm -f gpurun_out/prof_r2_$n.ncu-rep            # the reports are ~23 MB each: only the summaries travel back
done
ls -la gpurun_out | tail -8
