#!/bin/bash
# r28: full -m gpu suite (8-mer partition on packed words), bench (device + e2e), C4 / C5 shapes on the 3.1 Gbp index.
TAG=${1:-r28}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cut -c1-1500 gpurun_out/${TAG}_pytest.txt | tail -12
PREFIX=data/_gen/syn/syn3100
python bench.py --steps 5 --warmup 3 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"])
PY
python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 0 --ref-pb 0 --check 100 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err; cut -c1-760 gpurun_out/${TAG}_modes.jsonl
tail -5 gpurun_out/${TAG}_bench.err
