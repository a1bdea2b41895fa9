#!/bin/bash
# Every BASELINE config on one GPU through bench.py -> one JSON line each (profiles/r2_bench_all_configs.jsonl).
out="${1:-gpurun_out/r2_bench_all_configs.jsonl}"
: > "$out"
for args in "--workload cfg1 --precision fp32" "--workload cfg2 --precision fp32 --variant contrast" \
            "--workload cfg2 --precision fp32" "--workload cfg3 --precision fp32" \
            "--workload cfg3 --precision tf32 --variant contrast" "--workload cfg3 --precision tf32" \
            "--workload cfg4 --precision tf32 --steps 5"; do
  timeout 400 python bench.py --steps 20 --warmup 5 $args 2>>"${out%.jsonl}.err" | grep '^{' >> "$out"
done
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    r = d.get("roofline") or {}
    print(d["config"]["workload"][:34], d["run"]["precision"], d["run"]["path"], "us/step %.1f" % (d["ms_per_step"] * 1e3),
          "Mutt/s %.2f" % (d["value"] / 1e6), "launches/step", d["gpu_launches"] // d["steps"], "frac %.3f" % r.get("frac", 0),
          "e2e %.2f" % (d["e2e"]["value"] / 1e6), "cpu", d.get("cpu_baseline", {}).get("kind"), "%.0f" % d.get("cpu_baseline", {}).get("value", 0))
PY
