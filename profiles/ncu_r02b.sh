#!/bin/bash
# ncu captures after the table-driven exponential and the control/worker split (round 2, second half): the persistent fit
# and the streaming passes K2/K3 at 10^8 samples.  Each only after its own command exited 0 without ncu.
#   gpurun --timeout 1500 -- 'bash profiles/ncu_r02b.sh'
set -u
run() {  # name, kernel regex, skip, count, command...
    local name=$1 regex=$2 skip=$3 count=$4; shift 4
    "$@" > gpurun_out/plain_$name.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -f -k "regex:$regex" -s "$skip" -c "$count" -o gpurun_out/r02b_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name rc=$?"
}
run persist k_persistent_fit 2 1 python profiles/prof_kernels.py fit 1000000
run k2 "k_normal_eq_tma|k_cost_tma" 2 4 python profiles/prof_kernels.py k2 100000000
ls -la gpurun_out/r02b_*.ncu-rep
