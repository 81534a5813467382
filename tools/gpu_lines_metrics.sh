#!/bin/bash
# per-kernel time and DRAM bytes of one single-pair call (cost kernel + line kernels) at three settings -> gpurun_out/lines_ncu.txt
O=gpurun_out; mkdir -p $O
for cfg in "640x480 224 3" "1920x1080 128 1" "1280x720 64 0"; do
  set -- $cfg
  python tools/ab_variants.py --batch 1 --reps 1 --size $1 --disp $2 --mode $3 default > /dev/null 2>&1   # creates the cached inputs
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active \
    --clock-control none -k regex:"k_lines|k_wta_lines|k_cost" --launch-skip 6 -c 3 --csv --log-file $O/ln.csv \
    python tools/ab_variants.py --child --batch 1 --reps 1 --size $1 --disp $2 --mode $3 --data /tmp/ab_data_$1_$2.npz > /dev/null 2>&1
  python - "$1 D=$2 mode=$3" $O/ln.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(l for l in open(sys.argv[2]) if not l.startswith('=='))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[1:]:
    if len(r)>vi: print(sys.argv[1].ljust(26), r[ki].replace('void ','').replace('svgpu::<unnamed>::','')[:34].ljust(34), r[mi].ljust(58), r[vi], r[ui])
PY
done 2>&1 | tee $O/lines_ncu.txt
rm -f $O/ln.csv
