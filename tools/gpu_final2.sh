#!/bin/bash
mkdir -p gpurun_out
TAG=r01e
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --codec 4mz > gpurun_out/${TAG}_bench_4mz.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --codec 4mz --impl reference > gpurun_out/${TAG}_bench_4mz_reference.json 2>> gpurun_out/${TAG}_bench.err
for f in bench bench_4mz; do python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_$f.json")); print("$f value %.1f e2e %.2f" % (j["value"], j["e2e"]["value"]), j["detail"]["step_ms"], j["clocks"])
PY
done
timeout 600 python tools/latency_per_block.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -k "cli" 2>&1 | tail -2
