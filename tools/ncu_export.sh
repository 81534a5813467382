#!/bin/bash
# On the GPU box: turn an .ncu-rep into small text artefacts (raw metrics CSV + per-source-line and per-SASS CSVs, gzipped) and
# delete the report (gpurun only brings back 64 MiB).   usage: tools/ncu_export.sh <report.ncu-rep>
rep="$1"; base="${rep%.ncu-rep}"
ncu -i "$rep" --page raw --csv > "${base}_raw.csv" 2>/dev/null
ncu -i "$rep" --page source --csv --print-source cuda > "${base}_src_cuda.csv" 2>/dev/null
ncu -i "$rep" --page source --csv --print-source sass > "${base}_src_sass.csv" 2>/dev/null
gzip -f "${base}_src_cuda.csv" "${base}_src_sass.csv"
rm -f "$rep"
ls -la "${base}"_*
