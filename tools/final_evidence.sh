# Round-end evidence: headline benches, the ncu launch list and two `ncu --set full` captures (reports are
# converted to their raw CSV page on the box: a report with this library's SASS embedded is ~60 MB).
mkdir -p gpurun_out
timeout 400 python bench.py --steps 3 --warmup 3 --imad > gpurun_out/final_cfg4.json 2> gpurun_out/final.err; cut -c1-200 gpurun_out/final_cfg4.json
timeout 200 python bench.py --config cfg3 > gpurun_out/final_cfg3.json 2>> gpurun_out/final.err; cut -c1-150 gpurun_out/final_cfg3.json
timeout 200 python bench.py --config cfg2 > gpurun_out/final_cfg2.json 2>> gpurun_out/final.err; cut -c1-150 gpurun_out/final_cfg2.json
CMD="python bench.py --steps 2 --warmup 1 --batch 28 --e2e-batch 4 --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/final_b28.json 2>> gpurun_out/final.err && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_b28.csv $CMD > gpurun_out/ncu_a.log 2>&1
CMD2="python bench.py --steps 1 --warmup 1 --batch 14 --e2e-batch 2 --no-cpu-baseline --no-prof"
timeout 200 $CMD2 > gpurun_out/final_b14.json 2>> gpurun_out/final.err && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_pass -s 2 -c 2 -o /tmp/prof_final2_ks -f $CMD2 > gpurun_out/ncu_b.log 2>&1 && ncu -i /tmp/prof_final2_ks.ncu-rep --page raw --csv > gpurun_out/prof_final2_ks_raw.csv 2>/dev/null
CMD3="python bench.py --config cfg3 --batch 256 --e2e-batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-prof"
timeout 200 $CMD3 > gpurun_out/final_cfg3_b256.json 2>> gpurun_out/final.err && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_pass -s 2 -c 2 -o /tmp/prof_final2_cfg3_ks -f $CMD3 > gpurun_out/ncu_c.log 2>&1 && ncu -i /tmp/prof_final2_cfg3_ks.ncu-rep --page raw --csv > gpurun_out/prof_final2_cfg3_ks_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_b.log; tail -2 gpurun_out/ncu_c.log; tail -3 gpurun_out/final.err
