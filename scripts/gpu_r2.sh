#!/bin/bash
# Round-2 GPU pass: full parity suite, smoke, bench lines. usage (under gpurun): bash scripts/gpu_r2.sh <tag> [quick]
tag=${1:-r2}
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
if [ "$2" != "quick" ]; then
  python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -80 > gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
fi
python bench.py --steps 30 --warmup 5 --train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step %.3f serial %.3f fps %.0f e2e %.0f step_frac %.3f roof %.3f launches %d" % (d["ms_per_step"], d["serial_ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["step_frac"], d["roofline"]["frac"], d["gpu_launches"]))
print("   ", {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
print("   LN", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get("layernorm_f1") or {}).items() if k.endswith("_ms")})
print("   bf16", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get("bf16_canvas") or {}).items() if "ms" in k or "frames" in k})
print("   train", d.get("train"))
print("   cpu", d.get("cpu_baseline"))
PY
