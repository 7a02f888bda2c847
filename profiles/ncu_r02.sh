#!/bin/bash
# ncu captures of round 2 (one gpurun call, one B200): each kernel only after its own command exited 0 without ncu.
#   gpurun --timeout 1500 -- 'bash profiles/ncu_r02.sh'
set -u
run() {  # name, kernel regex, skip, count, command...
    local name=$1 regex=$2 skip=$3 count=$4; shift 4
    "$@" > gpurun_out/plain_$name.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c "$count" -o gpurun_out/r02_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name rc=$?"
}
run persist k_persistent_fit 2 1 python profiles/prof_kernels.py fit 1000000
run batch_exact k_batched_fit 1 1 python profiles/prof_kernels.py batch 65536 64 exact
run batch_fast k_batched_fit 1 1 python profiles/prof_kernels.py batch 65536 64
run gather "k_gather_samples|k_project|k_owner|k_scan" 6 6 python profiles/prof_kernels.py gather_scene
ls -la gpurun_out/*.ncu-rep | tail -6
