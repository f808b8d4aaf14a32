#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel regex> <launch count> [extra ncu args...] -- <command...>
# runs `ncu --set full`, exports the raw metric page as CSV (small) into gpurun_out/<name>.csv and drops the .ncu-rep
# (gpurun only copies 64 MiB back).  Read here with tools/ncu_table.py.
name=$1; regex=$2; count=$3; shift 3
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
mkdir -p gpurun_out
rep=/tmp/${name}.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -c "$count" "${extra[@]}" -f -o /tmp/${name} "$@" > gpurun_out/${name}.log 2>&1
ncu -i $rep --page raw --csv > gpurun_out/${name}.csv 2>> gpurun_out/${name}.log
ls -la $rep >> gpurun_out/${name}.log 2>&1
rm -f $rep
