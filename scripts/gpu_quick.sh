# quick A/B of the pipelined step: K2 register cap (96 = natural, 88, 80) x scatter CTAs per SM
run() { tag=$1; shift; python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-layernorm "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err || tail -5 gpurun_out/$tag.err; }
for regs in 88 80; do
  sed -i "s/^__global__ void __maxnreg__([0-9]*)\$/__global__ void __maxnreg__($regs)/" mask_bev_b200/csrc/pfn_tcw2.cuh
  python -m mask_bev_b200.build --force > /dev/null 2>&1
  for c in 1; do run r2c_regs${regs}_ctas$c --scatter-ctas $c; done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms/step %.3f serial %.3f fps %.0f e2e %.0f step_frac %.3f" % (d["ms_per_step"], d["serial_ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["step_frac"]))
        print("   ", {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    except Exception as e: print(f, "FAILED", e)
PY
