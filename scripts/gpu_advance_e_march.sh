#!/usr/bin/env bash
# Round 2, GPU call X (1 GPU): advance_e with threads marching up z -- parity (all variants), timing at 1024^3 against one voxel per thread
set -u
mkdir -p gpurun_out
S=gpurun_out/r2x_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_fields.py tests/test_gpu_aniso.py -q -m gpu -p no:cacheprovider --timeout=300 -rfEs > gpurun_out/r2x_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2x_pytest.log | tail -20 | tee -a $S
tail -60 gpurun_out/r2x_pytest.log > gpurun_out/r2x_pytest_tail.txt
for m in 0 1; do
  echo "fields.march_z=$m" | tee -a $S
  VPB_FIELDS_MARCH_Z=$m timeout 300 python scripts/fields_std.py 1024 2>&1 | tail -1 | tee -a $S
  VPB_FIELDS_MARCH_Z=$m timeout 300 python scripts/fields_std.py 512 2>&1 | tail -1 | tee -a $S
done
for m in 1 2; do
  VPB_FIELDS_MARCH_Z=$m timeout 400 python bench.py --field-cells 1024 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-deck-e2e > gpurun_out/r2x_bench_march$m.json 2> gpurun_out/r2x_bench_march$m.err
  echo "bench march=$m rc=$?" | tee -a $S
done
python - <<'PY' | tee -a $S
import json
for m in (1, 2):
    try:
        d = json.loads([l for l in open("gpurun_out/r2x_bench_march%d.json" % m) if l.startswith("{")][-1])
        f = d["fields_c2"]
        print("march", m, {k: (round(v["avg_launch_ms"], 3), round(v["frac"], 3)) for k, v in f.items() if isinstance(v, dict) and "frac" in v})
    except Exception as e:
        print("march", m, "failed", e)
PY
