#!/bin/bash
# r19: full -m gpu suite (packed entry point, rescue fast path, multi-device pool on one GPU, C3 with the sampled SA), bench at its
# default (C3: e2e through kb_map_chunk_packed, whole-program leg), CLI stage trace at C3 on 5 M reads, ncu of the rescue / candidate kernels.
TAG=${1:-r19}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt; nproc >> gpurun_out/${TAG}_gpu.txt; free -g | head -2 >> gpurun_out/${TAG}_gpu.txt
( time python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cat gpurun_out/${TAG}_pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 5500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
PREFIX=data/_gen/syn/syn3100
KART_B200_TRACE=1 python scripts/cli_compare.py --pairs 2500000 --prefix $PREFIX --error 0.01 --ours-only > gpurun_out/${TAG}_cli_c3.json 2> gpurun_out/${TAG}_cli_trace.txt; cat gpurun_out/${TAG}_cli_c3.json; grep -v "^\[kart trace\] \(read\|format\)" gpurun_out/${TAG}_cli_trace.txt | tail -30; grep "read \|format" gpurun_out/${TAG}_cli_trace.txt | tail -12
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_rescue_fast|k_cand_pair|k_unpack' -s 3 -c 3 -o gpurun_out/${TAG}c3_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}c3_ncu_full.log | cut -c1-200
tail -20 gpurun_out/${TAG}_bench.err
