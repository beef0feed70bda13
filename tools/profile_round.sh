#!/bin/bash
# One GPU call that collects the round's evidence into gpurun_out/ (copied to profiles/ afterwards):
#   tools/profile_round.sh TAG        e.g. r02c
# Every ncu pass runs after the same command has exited 0 without ncu.  The .ncu-rep files are condensed on the box
# (tools/ncu_summary.py, tools/sass_slots.py) and only the headline kernel's report travels back: gpurun_out/ is
# capped at 64 MiB.
T=${1:-r02}
O=gpurun_out
set -x
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
# launch list of the bench command (short form: same kernels, fewer steps)
BSHORT="python bench.py --steps 2 --warmup 3 --sustained-seconds 0 --no-cpu-baseline --no-other-configs"
$BSHORT > $O/${T}_bench_short.json 2>> $O/${T}_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $BSHORT > $O/${T}_ncu_launches.log 2>&1

capture() {   # capture NAME TAG ROWS VARIANT WARPS_PER_ROW SLOT_TAG
    local name=$1 tag=$2 rows=$3 var=$4 wpr=$5 slot=$6
    python tools/run_one.py $tag $rows $var 4 >> $O/${T}_run_one.log 2>&1 || return
    ncu --set full --clock-control none --import-source on --launch-skip 3 --launch-count 1 -k regex:polymul_kernel \
        -o /tmp/${T}_prof_$name -f python tools/run_one.py $tag $rows $var 4 >> $O/${T}_ncu_full.log 2>&1 || return
    python tools/ncu_summary.py /tmp/${T}_prof_$name.ncu-rep $O/${T}_ncu_$name.txt > /dev/null 2>> $O/${T}_ncu_full.log
    local vname=$(python - <<PY
import sys
sys.path.insert(0, "tiny-ntt_b200")
import tntt
from bench import PARAMS
p = PARAMS["$tag"]
pl = tntt.get_plan(p["n"], p["q"], p["psi"], True)
v = pl.default_variant if $var < 0 else $var
print(dict(pl.variants())[v].split(" ")[0])
PY
)
    python tools/sass_slots.py /tmp/${T}_prof_$name.ncu-rep --warps-per-row $wpr --tag $slot --variant $vname \
        --out $O/${T}_sass_hist_$vname.txt --json $O/${T}_sass_slots.json > /dev/null 2>> $O/${T}_ncu_full.log
}
V3=$(python - <<'PY'
import sys
sys.path.insert(0, "tiny-ntt_b200")
import tntt
p = tntt.get_plan(4096, (1 << 60) - (1 << 14) + 1, 431606828070683274, True)
print([v for v, d in p.variants() if d.startswith("u64_n12_r4_p1_a1_red3_b3_s1_t0")][0])
PY
)
VG=$(python - <<'PY'
import sys
sys.path.insert(0, "tiny-ntt_b200")
import tntt
p = tntt.get_plan(4096, (1 << 60) - (1 << 14) + 1, 431606828070683274, True)
print([v for v, d in p.variants() if d.startswith("u64_n12_r4_p1_a2_red1_b2_s0_t0_pad")][0])
PY
)
capture n4096_60 n4096_60 8192 -1 8 n4096_60
cp /tmp/${T}_prof_n4096_60.ncu-rep $O/                      # the one report that travels back
capture n4096_60_red1 n4096_60 8192 $VG 8 n4096_60_generic  # the default of every other 60-bit prime
capture n4096_60_red3 n4096_60 8192 $V3 8 n4096_60_barrett  # the reference's Barrett arithmetic in the same kernel
capture n4096_24 n4096_24 16384 -1 8 n4096_24
capture n1024_24 n1024_24 65536 -1 1 n1024_24
capture dilithium dilithium 262144 -1 0.5 dilithium
python tools/bench_variants.py n4096_60 > $O/${T}_variant_sweep.jsonl 2> $O/${T}_variant_sweep.err
python tools/rns_bench.py > $O/${T}_rns.jsonl 2> $O/${T}_rns.err
python tools/rns_bench.py --bits 23 --limbs 1,16 >> $O/${T}_rns.jsonl 2>> $O/${T}_rns.err
du -sh $O; ls -la $O | tail -30
