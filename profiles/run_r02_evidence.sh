#!/bin/bash
# Round-2 evidence run (one B200): tests, the bench line, the ncu launch list of the same command, one --set full
# capture of the dominant kernel, every BASELINE.json workload, and the one-frame launch lists.
# Everything lands in gpurun_out/; profiles/collect_r02.py turns it into the committed summaries.
o=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $o/r02_tests.log
python bench.py > $o/r02_bench.json 2> $o/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b2d|rows::' -c 600 --csv --log-file $o/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $o/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o $o/r02_rows_full \
    python profiles/rows_ab.py --one 64 2 > $o/r02_rows_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/r02_f1_launches.csv \
    python profiles/f1_latency.py waymo_test 3 > $o/r02_f1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/r02_f1_train_launches.csv \
    python profiles/f1_latency.py waymo_test 3 TRAIN > $o/r02_f1_train.log 2>&1
bash profiles/run_workloads.sh
cat $o/r02_tests.log
tail -c 600 $o/r02_bench.json
