#!/bin/bash
# per-kernel device times (ncu, cold-cache/serialised: compare shares) of the binning kernels of one fused cfg3 view
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tile_|preprocess_fwd|emit|radix|onesweep|finalize" -s ${1:-14} -c ${2:-14} --csv python tools/prof_pass.py cfg3 4 fused 2>/dev/null | python -c "
import csv, sys
for r in csv.reader(sys.stdin):
    if len(r) > 10 and r[0].isdigit():
        print(r[4][:70], r[-1])
"
