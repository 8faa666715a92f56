#!/bin/bash
# Run under gpurun: rebuild the chaos-game kernel with different register caps, time C2 and
# count executed instructions / local-memory requests of one launch.
for mb in ${@:-2 3 4}; do
  rm -f flam3_b200/lib/f3b_iterate.o
  make lib NVFLAGS_EXTRA="-DF3B_ITER_MIN_BLOCKS=$mb" > /dev/null 2>&1
  echo "== F3B_ITER_MIN_BLOCKS=$mb"
  for rep in 1 2; do
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4e  ms/step %.2f  iter kernel ms %.2f  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['clocks']))
    else: print(l.rstrip())
"
  done
  ncu --metrics smsp__inst_executed.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,sm__cycles_elapsed.max,gpu__time_duration.sum \
      --clock-control none -k regex:f3b_iterate -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "inst_executed|local_op|issue_active|warps_active|wavefronts|cycles|duration"
done
