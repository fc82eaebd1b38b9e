#!/bin/bash
# Call 40: the default bench line once more on the final tree (what the driver runs at round end).
O=gpurun_out/r02c40; mkdir -p $O
S=$(date +%s); timeout 105 python bench.py --steps 20 --warmup 5 > $O/bench.out 2> $O/bench.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
tail -1 $O/bench.out > $O/bench.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c40/bench.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "clocks", d.get("clocks"))
    print("roofline", d["roofline"]["frac"], "train", (d.get("train") or {}).get("ms_per_step"), "parity ok", (d.get("parity") or {}).get("ok"))
except Exception as e:
    print("no line:", e)
PY
