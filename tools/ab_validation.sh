# one-off A/B + validation batch (round 2): full GPU suite with and without PDL, then step-time variants
set +e
timeout 600 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2_ab_tests_default.log 2>&1; tail -2 gpurun_out/r2_ab_tests_default.log
ETB_PDL=1 timeout 600 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2_ab_tests_pdl.log 2>&1; tail -2 gpurun_out/r2_ab_tests_pdl.log
for v in "X=1" "ETB_CONV_2SM=3" "ETB_PDL=1" "ETB_PDL=1 ETB_CONV_2SM=3" "ETB_EPI_STAGE=0"; do
  env $v timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-e2e > gpurun_out/r2_bench_v.json 2> gpurun_out/r2_bench_v.err
  python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2_bench_v.json").read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
