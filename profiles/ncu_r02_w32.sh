# ncu capture of the persistent kernel running 2000 sweeps of 32 trial points (r02_w32.ncu-rep, profiles/r02_ncu_tables.md)
#   gpurun --timeout 900 -- "bash profiles/ncu_r02_w32.sh"   (after the same command has exited 0 without ncu)
set -e
cd /root/repo
BRDFGPU_SPEC_JAC=$((16 + 32*256)) ncu --set full --clock-control none --import-source on -k regex:k_persistent_fit -s 3 -c 1 -f -o gpurun_out/r02_w32 python profiles/persist_bench.py > gpurun_out/ncu_w32.log 2>&1 || tail -5 gpurun_out/ncu_w32.log
ls -la gpurun_out/r02_w32.ncu-rep
