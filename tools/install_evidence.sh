#!/bin/bash
# install_evidence.sh TAG: copy what tools/profile_round.sh TAG left in gpurun_out/ into profiles/ under the round's names
T=$1
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
cp $G/${T}_bench.json $P/r02_bench_n4096_60_1gpu.json
cp $G/${T}_bench_reference.json $P/r02_bench_reference_arm.json
cp $G/${T}_launches.csv $P/r02_launches_bench.csv
for f in $G/${T}_ncu_n4096_60.txt $G/${T}_ncu_n4096_60_red1.txt $G/${T}_ncu_n4096_60_red3.txt $G/${T}_ncu_n4096_24.txt $G/${T}_ncu_n1024_24.txt $G/${T}_ncu_dilithium.txt; do
    cp $f $P/$(basename $f | sed "s/${T}_/r02_/")
done
rm -f $P/r02_sass_hist_u32_*.txt $P/r02_sass_hist_u64_n12_r4_p1_a2_*.txt $P/r02_sass_hist_u64_n12_r4_p1_a1_red3*.txt
for f in $G/${T}_sass_hist_*.txt; do cp $f $P/$(basename $f | sed "s/${T}_/r02_/"); done
sed "s/${T}_sass_hist/r02_sass_hist/" $G/${T}_sass_slots.json > $P/sass_slots.json
cp $G/${T}_variant_sweep.jsonl $P/r02_variant_sweep.jsonl
cp $G/${T}_rns.jsonl $P/r02_rns.jsonl
for t in dilithium n1024_24 n4096_24; do [ -s $G/${T}_bench_$t.json ] && cp $G/${T}_bench_$t.json $P/r02_bench_${t}_1gpu.json; done
python - <<'PY'
import json, re
tj = json.load(open('profiles/traffic.json'))
sl = json.load(open('profiles/sass_slots.json'))
def grab(path):
    t = open(path).read()
    rd = float(re.search(r"dram__bytes_read.sum\s+([\d.]+) Mbyte", t).group(1)) * 1e6
    wr = float(re.search(r"dram__bytes_write.sum\s+([\d.]+) Mbyte", t).group(1)) * 1e6
    return rd, wr, int(re.search(r"launch__grid_size\s+(\d+)", t).group(1))
for tag, f, ppc in (("n4096_60", "r02_ncu_n4096_60.txt", 1), ("n4096_24", "r02_ncu_n4096_24.txt", 1), ("n1024_24", "r02_ncu_n1024_24.txt", 8), ("dilithium", "r02_ncu_dilithium.txt", 16)):
    rd, wr, grid = grab("profiles/" + f)
    rows = grid * ppc
    tj[tag] = {"dram_bytes_per_row": round((rd + wr) / rows), "rows_profiled": rows, "read_bytes": int(rd), "write_bytes": int(wr),
               "source": f"profiles/{f} (default variant {sl[tag]['variant']})"}
json.dump(tj, open('profiles/traffic.json', 'w'), indent=1)
for k, v in sl.items():
    print(k, v["variant"], v["inst_per_warp"], v["imad_wide_per_warp"], v["imad_narrow_per_warp"], v["pipe_cycles_per_warp"],
          "ceil %.2f M" % (148 * 4 * 1965e6 / (v["warps_per_row"] * v["pipe_cycles_per_warp"]) / 1e6))
d = json.load(open('profiles/r02_bench_n4096_60_1gpu.json'))
print("value", d["value"], d["sustained"]["value"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"])
for o in d["other_configs"]:
    print(o["tag"], o["rows_total"], "%.4g" % o["value"], o["kernel_variant"], "hbm %.3f" % o["hbm_frac"], o.get("mul_pipe_frac"))
PY
