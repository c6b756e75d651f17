#!/bin/bash
# End-of-round evidence on one B200 (gpurun -- bash tools/round_evidence.sh <tag>): bench lines first (no profiler), then the ncu
# launch list of the same bench command and full captures of the shipped kernels.  Everything lands in gpurun_out/.
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
timeout 600 python bench.py > $o/${tag}_bench_final.json 2> $o/${tag}_bench_final.err || echo "bench failed"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench_final.err || echo "reference arm failed"
timeout 300 python tools/time_configs.py > $o/${tag}_time_configs_final.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_ws|jacobi|scan_|cov|herk|rootmusic|find_local|calibrate" -c 400 --csv --log-file $o/${tag}_launches_raw.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-sustained > $o/${tag}_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_ws_kernel -s 4 -c 1 -o $o/${tag}_chain_ws -f \
    python tools/prof_chain.py > $o/${tag}_ncu_chain.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"cov16|jacobi_group|scan_tc" -s 9 -c 3 -o $o/${tag}_cfg5_kernels -f \
    python tools/prof_cfg5.py > $o/${tag}_ncu_cfg5.log 2>&1
tail -c 400 $o/${tag}_bench_final.err
python - <<PY
import json
d=json.loads(open("$o/${tag}_bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["sustained"]["ms_per_step"], d["e2e"]["value"])
for k,v in d["other_configs"].items(): print(k, v.get("ms", v.get("ms_per_step")), v.get("frames_per_s"))
PY
tail -n 6 $o/${tag}_time_configs_final.log
