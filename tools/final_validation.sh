set +e
timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2_final_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_final_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py --kernel-table gpurun_out/r2_step_kernel_table_final2.md > gpurun_out/r2_bench_ssod640_final2.json 2> gpurun_out/r2_bench_ssod640_final2.err; echo "rc=$?"
timeout 300 python bench.py --config sup32 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_sup32_final2.json 2> gpurun_out/r2_bench_sup32_final2.err
timeout 400 python bench.py --config ssod1280 --no-cpu-baseline > gpurun_out/r2_bench_ssod1280_final2.json 2> gpurun_out/r2_bench_ssod1280_final2.err
python - <<'PY'
import json
for f in ("r2_bench_ssod640_final2","r2_bench_sup32_final2","r2_bench_ssod1280_final2"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), (d.get("gpu_eager_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), d["roofline"]["frac"], d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 400 ncu --nvtx --nvtx-include etb_step --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_step_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-e2e --nvtx-step --allow-invalid > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_step_launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_fwd2_kernel -c 2 -o gpurun_out/r2_prof_conv_fwd2_raw_3x3_256_final -f python tools/conv_bench.py --batch 32 --modes raw --only "3x3 256->256" --iters 1 > gpurun_out/ncu_fwd_final.log 2>&1; ls -la gpurun_out/r2_prof_conv_fwd2_raw_3x3_256_final.ncu-rep
