MBEV_TC_DBG=8 timeout 120 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^chunk" | head -8
