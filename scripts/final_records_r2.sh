#!/bin/bash
# Final round-2 records on one B200 (run under gpurun): both bench arms, the launch list of the bench command, one
# --set full capture each of the headline kernel at full size, R-MFCC on the tcgen05 kernel and the tcgen05 FBANK kernel.
# Every ncu pass runs only after the same command has exited 0 without ncu.
set -x
O=gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err || exit 1
python bench.py > $O/r2f_bench.json 2> $O/r2f_bench.err || exit 1
python bench.py --steps 2 --warmup 1 --no-presets --no-cpu-baseline > $O/r2f_plain.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-presets --no-cpu-baseline > $O/r2f_under_ncu.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:srfe_mfcc_tc -s 1 -c 1 -o $O/prof_r2f_cmfcc \
    python bench.py --steps 2 --warmup 1 --no-presets --no-cpu-baseline > $O/r2f_ncu_cmfcc.log 2>&1
bash scripts/profile_summary.sh $O/prof_r2f_cmfcc.ncu-rep $O/r2f_ncu_bench_cmfcc_tc_262144.txt; rm -f $O/prof_r2f_cmfcc.ncu-rep
python scripts/quick_bench.py 16384 R-MFCC,R-FBANK,C-FBANK > $O/r2f_quick.log 2>&1 || exit 1
python scripts/quick_bench.py 16384 R-FBANK,C-FBANK fbank_tc=2 >> $O/r2f_quick.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:srfe_mfcc_tc -s 3 -c 1 -o $O/prof_r2f_rmfcc \
    python scripts/quick_bench.py 16384 R-MFCC > $O/r2f_ncu_rmfcc.log 2>&1
bash scripts/profile_summary.sh $O/prof_r2f_rmfcc.ncu-rep $O/r2f_ncu_r_mfcc_tc_16384.txt; rm -f $O/prof_r2f_rmfcc.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:srfe_fbank_tc -s 3 -c 1 -o $O/prof_r2f_fbank \
    python scripts/quick_bench.py 16384 R-FBANK fbank_tc=2 > $O/r2f_ncu_fbank.log 2>&1
bash scripts/profile_summary.sh $O/prof_r2f_fbank.ncu-rep $O/r2f_ncu_r_fbank_tc_16384.txt; rm -f $O/prof_r2f_fbank.ncu-rep
cat $O/r2f_quick.log
head -c 600 $O/r2f_bench.json; echo
head -c 400 $O/r2f_bench_ref.json; echo
