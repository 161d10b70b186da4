#!/bin/bash
# last call of the round (1 GPU): what the driver runs at round end -- the GPU suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02z_final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02z_final_smoke.txt
python bench.py > gpurun_out/r02z_final_bench.json 2> gpurun_out/r02z_final_bench.err; tail -2 gpurun_out/r02z_final_bench.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02z_final_bench.json"))
print("value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 2), {k: round(j["e2e"][k], 1) for k in ("compress_GBps", "decompress_GBps")},
      "roofline", j["roofline"]["kernel"], round(j["roofline"]["frac"], 4), "cpu", j["cpu_baseline"] and round(j["cpu_baseline"]["value"], 2), "launches", j["gpu_launches"], j["clocks"])
PY
