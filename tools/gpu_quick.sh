#!/bin/bash
# quick iteration: kNN/graph parity tests + one short bench (no CPU baseline, no coclust)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cgraph.py -m gpu -q -x --timeout=200 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-coclust > gpurun_out/bench_quick.log 2>&1
echo "bench exit $?"
python - <<'PY'
import json
for ln in open('gpurun_out/bench_quick.log'):
    if ln.startswith('{"metric"'):
        d=json.loads(ln)
        print("cells/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]))
        print("stages", d["stages_ms"])
        print("roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","issued_tflops","ms_per_launch","share_of_step")})
        print("filter", d["config"]["knn_filter"], "clocks", d["clocks"])
        break
else:
    print(open('gpurun_out/bench_quick.log').read()[-3000:])
PY
