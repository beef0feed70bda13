#!/bin/bash
# One GPU call that collects the round's evidence into gpurun_out/ (copied to profiles/ afterwards):
#   tools/profile_round.sh TAG        e.g. r02b
# Every ncu pass runs after the same command has exited 0 without ncu.
T=${1:-r02}
O=gpurun_out
set -x
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
# launch list of the bench command (short form: same kernels, fewer steps)
BSHORT="python bench.py --steps 2 --warmup 3 --sustained-seconds 0 --no-cpu-baseline --no-other-configs"
$BSHORT > $O/${T}_bench_short.json 2>> $O/${T}_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $BSHORT > $O/${T}_ncu_launches.log 2>&1
# full capture of the headline kernel (8192 rows, 4th launch), with source counters for the executed-instruction mix
python tools/run_one.py n4096_60 8192 -1 4 > $O/${T}_run_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on --launch-skip 3 --launch-count 1 -k regex:polymul_kernel \
    -o $O/${T}_prof_n4096_60 -f python tools/run_one.py n4096_60 8192 -1 4 > $O/${T}_ncu_full.log 2>&1
for cfg in "n4096_24 16384" "n1024_24 65536" "dilithium 262144"; do
    set -- $cfg
    python tools/run_one.py $1 $2 -1 4 >> $O/${T}_run_one.log 2>&1 && \
    ncu --set full --clock-control none --import-source on --launch-skip 3 --launch-count 1 -k regex:polymul_kernel \
        -o $O/${T}_prof_$1 -f python tools/run_one.py $1 $2 -1 4 >> $O/${T}_ncu_full.log 2>&1
done
python tools/bench_variants.py n4096_60 > $O/${T}_variant_sweep.jsonl 2> $O/${T}_variant_sweep.err
python tools/rns_bench.py > $O/${T}_rns.jsonl 2> $O/${T}_rns.err
python tools/rns_bench.py --bits 23 --limbs 1,16 >> $O/${T}_rns.jsonl 2>> $O/${T}_rns.err
ls -la $O | tail -20
