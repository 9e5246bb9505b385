#!/bin/bash
# Round-2 GPU pass: full parity suite, bench lines (pipelined step with the three K3 choices), smoke.
# usage (from the repo root, under gpurun): bash scripts/gpu_r2.sh <tag> [quick]
tag=${1:-r2}
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
if [ "$2" != "quick" ]; then
  python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -80 > gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
for c in 0 2; do
  python bench.py --steps 30 --warmup 5 --scatter-ctas $c --no-cpu-baseline --no-layernorm > gpurun_out/${tag}_bench_ctas$c.json 2> gpurun_out/${tag}_bench_ctas$c.err || tail -5 gpurun_out/${tag}_bench_ctas$c.err
done
python - <<PY
import json
for f in ("${tag}_bench", "${tag}_bench_ctas0", "${tag}_bench_ctas2"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms/step %.3f serial %.3f fps %.0f e2e %.0f step_frac %.3f" % (d["ms_per_step"], d["serial_ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["step_frac"]))
        print("   ", {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
        if d.get("layernorm_f1"): print("   LN", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["layernorm_f1"].items() if k.endswith("_ms")})
    except Exception as e:
        print(f, "FAILED", e)
PY
