#!/bin/bash
# r40: the bench's whole-program leg with the CLI stage trace (where did 7 s go in r36?), ours run twice.
TAG=${1:-r40}
mkdir -p gpurun_out
KART_B200_TRACE=1 python bench.py --cpu-sample-pairs 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms  e2e %.3f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"])); print(d.get("e2e_program"))
PY
grep "kart trace" gpurun_out/${TAG}_bench.err | grep -v "read \|format\|map_batch\|pin " | tail -30
