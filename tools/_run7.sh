python tools/gpu_stage2_quick.py 2>&1 | tail -1 | tee gpurun_out/stage2_quick_pf.log
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_pf.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_pf.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sustained", d["sustained"]["value"], "C5", d["configs"]["C5"]["ms"], "C3", d["configs"]["C3"]["ms"])
PY
