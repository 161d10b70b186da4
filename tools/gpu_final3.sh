#!/bin/bash
mkdir -p gpurun_out
TAG=r01f
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --codec 4mz > gpurun_out/${TAG}_bench_4mz.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --codec 4mz --impl reference > gpurun_out/${TAG}_bench_4mz_reference.json 2>> gpurun_out/${TAG}_bench.err
for f in bench bench_4mz; do python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_$f.json")); print("$f value %.1f e2e %.2f" % (j["value"], j["e2e"]["value"]), j["detail"]["step_ms"], {k: round(v, 1) for k, v in j["e2e"].items() if k.endswith("GBps")})
PY
done
bash tools/gpu_hotlines.sh $TAG
timeout 600 python tools/quick_levels.py 4 2>&1 | grep level
