set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_pfn_tcw2" -s 3 -c 1 -f -o gpurun_out/tc1_pfn python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/tc1_ncu.log 2>&1
tail -3 gpurun_out/tc1_ncu.log
