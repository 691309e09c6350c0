#!/usr/bin/env bash
# first GPU pass: environment, parity tests, smoke, small + full bench, ncu launch list + one full capture
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; lscpu | grep -E 'Model name|Socket|Core|Thread'; } > $O/env.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 > $O/bench_128.json 2> $O/bench_128.err; echo "exit $?" >> $O/bench_128.err
for dep in 0 1; do
  VPB_ADVANCE_P_DEPOSIT=$dep timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_128_dep$dep.json 2> $O/bench_128_dep$dep.err
done
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_256.json 2> $O/bench_256.err; echo "exit $?" >> $O/bench_256.err
# ncu only after the same command exited 0 without it
if timeout 300 python bench.py --cells 128 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1.csv \
      python bench.py --cells 128 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 2 -c 2 -o $O/prof_advance_p_r1 \
      python bench.py --cells 128 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
fi
ls -la $O
