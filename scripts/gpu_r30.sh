#!/bin/bash
# r30: full -m gpu suite with the adaptive seeding quorums, bench (device + e2e), C4 / C5 shapes, and an A/B of k_align_part's warps / pool on C5.
TAG=${1:-r30}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cut -c1-1500 gpurun_out/${TAG}_pytest.txt | tail -8
PREFIX=data/_gen/syn/syn3100
python bench.py --steps 5 --warmup 3 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"])
PY
python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 0 --ref-pb 0 --check 100 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err; cut -c1-760 gpurun_out/${TAG}_modes.jsonl
for v in "KB_PART_WARPS=7104 KB_PART_POOL=3072" "KB_PART_WARPS=5920 KB_PART_POOL=2048" "KB_PART_WARPS=7104 KB_PART_POOL=2048" "KB_PART_WARPS=4736 KB_PART_POOL=4096"; do
  echo "== $v" >> gpurun_out/${TAG}_part_ab.jsonl
  env $v python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 0 --ref-pb 0 --check 0 >> gpurun_out/${TAG}_part_ab.jsonl 2>> gpurun_out/${TAG}_bench.err
done
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_part_ab.jsonl"):
    if ln.startswith("=="): print(ln.strip()); continue
    d = json.loads(ln); print(" ", d["mode"], round(d["device_ms"], 2), "align", round(d["stage_ms"]["align"], 2), "nw", round(d["stage_ms"]["nw"], 2))
PY
tail -5 gpurun_out/${TAG}_bench.err
