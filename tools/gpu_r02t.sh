#!/bin/bash
# r02t (1 GPU): E1 pass 1 with four positions per slice word; GPU suite; bench line of config 1 (e2e included, CPU baseline skipped)
mkdir -p gpurun_out
timeout 600 python tools/quick_decode.py 16 2 2>&1 | grep -v "^$" | tee gpurun_out/r02t_timing.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02t_pytest.txt
cat gpurun_out/r02t_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02t_bench.json"))
k = j["detail"]["kernel_ms_rank0"]; s = j["steps"]
print("value", round(j["value"], 1), "ms/step", round(j["ms_per_step"], 1), "e2e", {a: round(b, 1) for a, b in j["e2e"].items() if isinstance(b, (int, float))})
print({n: round(v["total_ms"] / s, 2) for n, v in k.items() if v["total_ms"] / s > 0.5})
print("roofline", j["roofline"])
PY
tail -3 gpurun_out/r02t_bench.err
