#!/bin/bash
OUT=gpurun_out/${1:-v3ncu}
mkdir -p $OUT
export ETCH_B200_INTER=v3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inter_conv_v3 -c 3 -f -o $OUT/full_v3 python tools/inter_v3_check.py time 8 5000 > $OUT/ncu.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu.log
ncu -i $OUT/full_v3.ncu-rep --page raw --csv > $OUT/full_v3.raw.csv 2>/dev/null
ls -la $OUT
