set -u
mkdir -p gpurun_out
S=gpurun_out/r2al_summary.txt
: > $S
timeout 170 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2al_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $S
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 -rfEs -x > gpurun_out/r2al_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2al_pytest.log | tail -10 | tee -a $S
