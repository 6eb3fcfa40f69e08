python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_late.log
python bench.py --steps 20 --warmup 3 --no-configs 2>&1 | tail -1 > gpurun_out/bench_1chunk.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_1chunk.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sustained", d["sustained"]["value"])
PY
