#!/bin/bash
# Final measurements of round 2 with the final code: headline (as the driver runs it), coupled with the per-problem split, C1 lanes, static, acoustic.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02_final_headline.err | tail -1 > gpurun_out/r02_final_headline.json
timeout 300 python bench.py --workload coupled --steps 2 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_final_coupled.json
timeout 300 python bench.py --workload c1 --steps 2 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_final_c1.json
timeout 200 python bench.py --workload static --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_final_static.json
timeout 200 python bench.py --workload acoustic --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_final_acoustic.json
python - <<'PY'
import json
for f in ("headline", "coupled", "c1", "static", "acoustic"):
    try:
        d = json.loads(open("gpurun_out/r02_final_%s.json" % f).read())
        print(f, d.get("value"), d.get("unit"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "roof", (d.get("roofline") or {}).get("frac"), d.get("clocks", {}).get("reasons"))
        if f == "coupled": print("   ", d["assembly"])
        if f == "headline": print("   ", d.get("assembly", {}).get("ms_regular"), d.get("lu", {}).get("ms_per_frequency"))
    except Exception as e:
        print(f, "FAILED", e)
PY
