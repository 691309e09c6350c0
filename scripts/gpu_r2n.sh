#!/usr/bin/env bash
# Round 2, GPU call N (1 GPU): configs[2] with the library's share of the trecon-part deck run (VPB_TRACE), full-sector
# interpolator stores
set -u
mkdir -p gpurun_out
S=gpurun_out/r2n_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_particles.py tests/test_gpu_fields.py tests/test_gpu_step.py tests/test_gpu_history.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs > gpurun_out/r2n_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2n_pytest.log | tail -10 | tee -a $S
timeout 300 python bench.py --no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err
echo "bench rc=$?" | tee -a $S
timeout 1200 python bench.py --workload harris --trecon-deck --no-e2e --no-cpu-baseline --no-deck-e2e --steps 20 --warmup 3 > gpurun_out/r2n_bench_harris.json 2> gpurun_out/r2n_bench_harris.err
echo "bench harris + trecon deck rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2n_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac", (d.get("roofline") or {}).get("frac"),
              "breakdown", {k: round(v, 3) for k, v in d["breakdown_ms_per_step"].items()})
        for k in ("trecon_deck", "trecon_deck_scaled"):
            if k in d: print("   ", k, json.dumps(d[k])[:1500])
    except Exception as e:
        print(f, "failed", e)
PY
