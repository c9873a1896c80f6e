#!/bin/bash
# C3/C4/C5 at genome scale ON THE GPU BOX: the index cannot travel (512 MiB snapshot limit), so it is made there: a seeded synthetic
# genome (scripts/make_syn_index.py) indexed with this repo's own builder (`kart index`, byte-identical to the reference's
# files, minutes instead of hours), then parity against the oracle, bench (C3 shape), the C4 / C5 shapes and the whole-program
# comparison with the reference. Not yet run on a GPU box (written at the end of round 1); every part is exercised elsewhere.
# Usage: bash scripts/gpu_c3.sh [Mbp=3100] [contigs=24] [tag=c3]      needs ~70 GB of host memory at 3100 Mbp (64-bit builder)
MBP=${1:-3100}; CONTIGS=${2:-24}; TAG=${3:-c3}
mkdir -p gpurun_out
# the 64-bit builder holds 2G x 8 B of suffix positions (+ text, BWT, histograms): refuse sizes the box cannot hold instead of
# driving it out of memory
AVAIL_GB=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
NEED_GB=$(( MBP * 23 / 1000 + 4 ))
if [ "$AVAIL_GB" -lt "$NEED_GB" ]; then echo "host has ${AVAIL_GB} GB available, ${MBP} Mbp needs ~${NEED_GB} GB: falling back to 2000 Mbp (32-bit builder)"; MBP=2000; fi
PREFIX=data/_gen/syn/syn$MBP
free -g | head -2 > gpurun_out/${TAG}_host.txt; nproc >> gpurun_out/${TAG}_host.txt
if [ ! -f $PREFIX.bwt ]; then
  ( time KART_INDEX_BUILDER=ours python scripts/make_syn_index.py $MBP $CONTIGS 12345 ) > gpurun_out/${TAG}_index_build.txt 2>&1
  tail -5 gpurun_out/${TAG}_index_build.txt
fi
ls -la data/_gen/syn/ >> gpurun_out/${TAG}_index_build.txt
# C3 shape: paired 2x150 @ 1 %, device-resident and end to end, reference kart -t N on a sample of the same reads
python bench.py --steps 5 --warmup 3 --prefix $PREFIX --error 0.01 --cpu-sample-pairs 000 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
# C4 / C5 shapes with 200 reads of each checked against the oracle
python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 50000 --ref-pb 500 --check 200 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_modes.jsonl
# whole program against the reference (sorted SAM vs kart -t N; kart -t 1 on 2 M reads of a 3.1 Gbp index takes minutes: --t1 only when asked)
python scripts/cli_compare.py --pairs 200000 --prefix $PREFIX --error 0.01 ${C3_T1:+--t1} --diff-out gpurun_out/${TAG}_cli_diff.txt > gpurun_out/${TAG}_cli.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_cli.json
# where the time goes at this index size
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --prefix $PREFIX --error 0.01 > gpurun_out/${TAG}_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_rescue|k_cand_pair|k_segments$|k_align_part' -s 7 -c 7 -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --prefix $PREFIX --error 0.01 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
tail -10 gpurun_out/${TAG}_bench.err
