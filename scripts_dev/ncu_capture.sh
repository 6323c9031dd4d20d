#!/bin/bash
# ncu captures of the round's final kernels (one GPU): launch list + full sets (raw + source pages as CSV) for ADMM and FISTA.
#   scripts_dev/ncu_capture.sh TAG     -> gpurun_out/TAG_launches.csv, TAG_ncu_full_{admm,fista}_{raw,src}.csv
tag=$1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --configs none --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_admm_cols" -s 12 -c 3 -f -o /tmp/admm \
    python bench.py --configs none --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/admm.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_admm_raw.csv
ncu -i /tmp/admm.ncu-rep --page source --csv > gpurun_out/${tag}_ncu_full_admm_src.csv
ncu --set full --clock-control none --import-source on -k regex:k_gd -s 12 -c 4 -f -o /tmp/fista \
    python bench.py --configs none --algo fista --n-iter 20 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/fista.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_fista_raw.csv
ncu -i /tmp/fista.ncu-rep --page source --csv > gpurun_out/${tag}_ncu_full_fista_src.csv
ls -la gpurun_out/${tag}_*
