#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01j}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --codec 4mz > gpurun_out/${TAG}_bench_4mz.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --codec 4mz --impl reference > gpurun_out/${TAG}_bench_4mz_reference.json 2>> gpurun_out/${TAG}_bench.err
[ -n "$SKIP12" ] || python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_bench_12steps.json 2>> gpurun_out/${TAG}_bench.err
for f in bench bench_4mz; do python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_$f.json")); print("$f value %.1f" % j["value"], "e2e", j["e2e"] and {k: round(v, 1) for k, v in j["e2e"].items() if k.endswith("GBps") or k == "value"}, j["detail"]["step_ms"])
PY
done
