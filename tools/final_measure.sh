#!/bin/bash
# Round-end measurement on the GPU box (one GPU): driver-style bench line, then -- each only after the plain run exited 0 -- the ncu launch
# list of the same command and one `ncu --set full` capture of a cfg2 pass, a cfg3 pass and an encode.  usage: tools/final_measure.sh <tag>
T=${1:-r2b}
O=gpurun_out
python bench.py > $O/${T}_bench_full.json 2> $O/${T}_bench_full.err || exit 1
python bench.py --steps 2 --warmup 1 --no-configs --no-e2e --no-cpu > $O/${T}_b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-e2e --no-cpu > $O/${T}_ncu_launch.log 2>&1
python tools/run_cfg.py cfg2 3 > $O/${T}_cfg2_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 20 -c 12 -o /tmp/${T}_cfg2 python tools/run_cfg.py cfg2 3 > $O/${T}_cfg2_ncu.log 2>&1
ncu -i /tmp/${T}_cfg2.ncu-rep --page raw --csv > $O/${T}_cfg2_raw.csv 2>/dev/null
ncu -i /tmp/${T}_cfg2.ncu-rep --page source --csv -k regex:k_decode > $O/${T}_cfg2_decode_src.csv 2>/dev/null
python tools/enc_profile.py cfg2 > $O/${T}_enc_plain.log 2>&1 || exit 1
python tools/enc_profile.py cfg3 >> $O/${T}_enc_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_enc --launch-skip 6 -c 3 -o /tmp/${T}_enc python tools/enc_profile.py cfg2 > $O/${T}_enc_ncu.log 2>&1
ncu -i /tmp/${T}_enc.ncu-rep --page raw --csv > $O/${T}_enc_raw.csv 2>/dev/null
ncu -i /tmp/${T}_enc.ncu-rep --page source --csv -k regex:k_enc_plan > $O/${T}_enc_plan_src.csv 2>/dev/null
tail -c 600 $O/${T}_bench_full.json; cat $O/${T}_enc_plain.log
