#!/bin/bash
# ncu captures at the round's final HEAD of the launches the earlier capture scripts do not cover: the classic log-filterbank
# kernel after the register-overlap change (R-FBANK, C-FBANK) and the two log-spectrogram FT-layout launches.  Run under gpurun;
# every ncu pass follows a plain run of the same command that exited 0.
set -x
O=gpurun_out
python scripts/quick_bench.py 16384 R-FBANK,C-FBANK,R-SPEC:ft,C-SPEC:ft > $O/r2h_plain.log 2>&1 || exit 1
for spec in "R-FBANK r_fbank" "C-FBANK c_fbank" "R-SPEC:ft r_spec_ft" "C-SPEC:ft cspec_ft"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:srfe_ -s 3 -c 1 -o $O/prof_r2h_$2 python scripts/quick_bench.py 16384 $1 > $O/r2h_ncu_$2.log 2>&1
  bash scripts/profile_summary.sh $O/prof_r2h_$2.ncu-rep $O/r2h_ncu_$2_16384.txt
  rm -f $O/prof_r2h_$2.ncu-rep
done
cat $O/r2h_plain.log
