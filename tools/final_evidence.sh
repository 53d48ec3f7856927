# Round-end evidence (round 2): headline benches, the ncu launch list of the same command, and `ncu --set full` captures of
# the key-switch kernels (cfg4) and of the 32-bit transform passes; reports are converted to their raw CSV page on the box
# (a report with this library's SASS embedded is ~60 MB).  Bench numbers are never taken under a profiler: every ncu
# command is preceded by the plain run of the same command.
mkdir -p gpurun_out
R=${R:-r02}
timeout 300 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_final_cfg4.json 2> gpurun_out/${R}_final.err; cut -c1-200 gpurun_out/${R}_final_cfg4.json
timeout 300 python bench.py --config cfg3 --steps 5 --warmup 3 > gpurun_out/${R}_final_cfg3.json 2>> gpurun_out/${R}_final.err; cut -c1-150 gpurun_out/${R}_final_cfg3.json
timeout 300 python bench.py --config cfg2 --steps 5 --warmup 3 > gpurun_out/${R}_final_cfg2.json 2>> gpurun_out/${R}_final.err; cut -c1-150 gpurun_out/${R}_final_cfg2.json
timeout 120 python bench.py --config cfg1 --steps 20 --warmup 5 > gpurun_out/${R}_final_cfg1.json 2>> gpurun_out/${R}_final.err; cut -c1-150 gpurun_out/${R}_final_cfg1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_final_reference_arm.json 2>> gpurun_out/${R}_final.err
CMD="python bench.py --steps 2 --warmup 1 --batch 56 --e2e-batch 4 --no-cpu-baseline --no-ntt --no-chain"
timeout 200 $CMD > gpurun_out/${R}_final_b56.json 2>> gpurun_out/${R}_final.err && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_b56.csv $CMD > gpurun_out/ncu_a.log 2>&1
CMD2="python bench.py --steps 1 --warmup 1 --batch 28 --e2e-batch 2 --no-cpu-baseline --no-prof --no-ntt --no-chain"
timeout 200 $CMD2 > gpurun_out/${R}_final_b28.json 2>> gpurun_out/${R}_final.err && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_pass -s 2 -c 2 -o /tmp/prof_${R}_ks -f $CMD2 > gpurun_out/ncu_b.log 2>&1 && ncu -i /tmp/prof_${R}_ks.ncu-rep --page raw --csv > gpurun_out/${R}_prof_ks_raw.csv 2>/dev/null
CMD3="python tools/ntt_probe.py 30 16 24 85"
timeout 200 $CMD3 > gpurun_out/${R}_ntt_probe.log 2>> gpurun_out/${R}_final.err && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 4 -c 4 -o /tmp/prof_${R}_ntt30 -f $CMD3 > gpurun_out/ncu_c.log 2>&1 && ncu -i /tmp/prof_${R}_ntt30.ncu-rep --page raw --csv > gpurun_out/${R}_prof_ntt30_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_b.log; tail -2 gpurun_out/ncu_c.log; tail -3 gpurun_out/${R}_final.err
