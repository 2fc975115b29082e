#!/bin/bash
# last-tree evidence on one B200: smoke, full -m gpu suite, the bench line (default flags), the reference arm, ncu launch list
# + full captures of the heaviest kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $OUT/r03_final_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | grep -v Warning | tail -6 | tee $OUT/r03_final_pytest.log
timeout 900 python bench.py > $OUT/r03_final_bench.json 2> $OUT/r03_final_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r03_final_bench.json"))
print("ms/step=%.4f rays/s=%.4e e2e=%.4e launches=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("gpu_launches")))
print(d.get("kernel_us")); print(d.get("roofline")); print(d.get("cpu_baseline")); print(d.get("clocks"))
P
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r03_final_bench_reference.json 2> $OUT/r03_final_bench_reference.err; echo "reference rc=$?"; cut -c1-600 $OUT/r03_final_bench_reference.json
bash profiles/run_ncu.sh r03 > $OUT/r03_run_ncu.log 2>&1; tail -3 $OUT/r03_run_ncu.log
