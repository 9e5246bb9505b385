mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_pfn_tcw2" -s 3 -c 1 -f -o gpurun_out/cv_pfn python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/cv_ncu.log 2>&1
tail -3 gpurun_out/cv_ncu.log
