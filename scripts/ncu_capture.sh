#!/bin/bash
# One `ncu --set full` capture of ONE launch on the GPU box, exported as the two CSV pages the repo's tools read
# (gpurun_out/ only carries 64 MiB back, a .ncu-rep with sources is larger).
# usage: ncu_capture.sh <name> <kernel regex> <launch-skip> <command...>
name=$1; regex=$2; skip=$3; shift 3
out=gpurun_out
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip $skip -c 1 -f -o /tmp/$name "$@" > $out/${name}_ncu.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>> $out/${name}_ncu.log
ncu -i /tmp/$name.ncu-rep --page source --csv > $out/${name}_src.csv 2>> $out/${name}_ncu.log
ls -la $out/${name}_*
