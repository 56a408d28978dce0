#!/bin/bash
# ncu_export.sh <report.ncu-rep>: keep the small exports of a capture (raw metrics CSV, gzipped source/SASS CSV) and drop the report itself
# (gpurun copies back at most 64 MiB per call; one --import-source report of k_step is ~25 MB).
rep=$1
base=${rep%.ncu-rep}
ncu -i $rep --page raw --csv > ${base}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > ${base}_source.csv.gz
rm -f $rep
ls -la ${base}_raw.csv ${base}_source.csv.gz | awk '{print $5, $9}'
