#!/bin/bash
# e2e (host in -> host out) of the C2 build against the number of staging threads.  usage: tools/e2e_threads.sh
nproc
for t in 4 8 12 16 24; do
  MCB_HOST_THREADS=$t python bench.py --no-cpu --no-coclust --no-parity --steps 3 --warmup 2 2>/dev/null > /tmp/e2e_$t.json
  python - $t <<'PY'
import json, sys
t = sys.argv[1]
d = json.load(open(f"/tmp/e2e_{t}.json"))
print("threads", t, "e2e ms", round(d["e2e"]["ms_per_step"], 2), "h2d", d["e2e"]["stages_ms"]["h2d_csc"], "d2h", d["e2e"]["stages_ms"]["d2h_edges"], "step", round(d["ms_per_step"], 2))
PY
done
