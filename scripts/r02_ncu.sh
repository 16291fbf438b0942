#!/bin/bash
# Round-2 ncu evidence (run under gpurun, ONE GPU): launch list of one eager OC20 step + `--set full` captures of the
# dominant GEMM and of the kernels this round changed.  Numbers printed by runs under ncu are never bench values.
OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
SHORT="python bench.py --layers 2 --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > $OUT/r02_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/r02_ncu_plain.log; exit 1; }
# one full eager step: skip the three warm-up steps (~2 700 launches each incl. torch kernels), record ~1.2 steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 3400 --csv --log-file $OUT/r02_ncu_launches.csv $CMD > $OUT/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
$SHORT > $OUT/r02_ncu_plain_short.log 2>&1 || { echo "short plain run failed"; exit 1; }
for spec in "gemm_f16_kernel:gemm:40:3" "gather_rotate_dx_pipe_kernel:grdx:2:1" "rotinv_reduce_fwd_pipe_kernel:rirfwd:2:1" ; do
  IFS=: read pat tag skip cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o $OUT/r02_ncu_$tag $SHORT > $OUT/r02_ncu_$tag.log 2>&1
  echo "$tag rc=$?"
done
ls -la $OUT/*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
