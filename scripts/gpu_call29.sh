#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py -m gpu -q -x -n 4 -p no:cacheprovider -k "sort" > $O/pytest_gpu29_sort.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu29_sort.log
tail -2 $O/pytest_gpu29_sort.log
timeout 1200 python bench.py > $O/b29_default.json 2> $O/b29_default.err; echo "exit $?" >> $O/b29_default.err
for si in 10 5; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 --sort-interval $si > $O/b29_si$si.json 2> $O/b29_si$si.err
done
python - <<'PY'
import json
for f in ["default","si10","si5"]:
    d=json.loads(open("gpurun_out/b29_%s.json"%f).read().strip().splitlines()[-1])
    l=d["advance_p_ms_by_launch"]
    print(f, "value %.3e ms/step %.2f avg %.2f frac %.3f fresh %.2f last %.2f sort %.2f clk %s"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"]))
    if "e2e" in d: print("  e2e", d["e2e"], "cpu", d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/b29_ref.json 2> $O/b29_ref.err; tail -c 700 $O/b29_ref.json
