for l in default _ab/nopdl.so _ab/nofused.so _ab/neither.so; do
  if [ $l = default ]; then python profiles/f1_latency.py waymo_test 30; else B2D_LIB_PATH=$PWD/$l python profiles/f1_latency.py waymo_test 30; fi
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f1_launches_new.csv python profiles/f1_latency.py waymo_test 2 > /dev/null 2>&1
