#!/bin/sh
# Export the raw metric page of an .ncu-rep (brought back from the GPU box in gpurun_out/) as CSV under profiles/ - the file
# bench.py reads `roofline.traffic` from.  Usage: tools/ncu_export.sh gpurun_out/msm_acc_r02.ncu-rep profiles/ncu_msm_accumulate_r02_raw.csv
set -e
ncu -i "$1" --page raw --csv > "$2"
echo "wrote $2 ($(wc -l < "$2") lines)"
