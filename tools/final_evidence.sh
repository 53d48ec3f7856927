# Round-end evidence (round 2): headline benches, the ncu launch list of the same command, and `ncu --set full` captures of
# the key-switch kernels of the auxiliary-basis path (cfg4); reports are converted to their raw CSV page on the box
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
timeout 600 python bench.py --steps 5 --warmup 3 --no-ntt --no-chain --no-single-thread --ks-aux 0 > gpurun_out/${R}_final_cfg4_ks_aux0.json 2>> gpurun_out/${R}_final.err; cut -c1-120 gpurun_out/${R}_final_cfg4_ks_aux0.json
CMD="python bench.py --steps 2 --warmup 1 --batch 64 --e2e-batch 4 --no-cpu-baseline --no-ntt --no-chain"
timeout 200 $CMD > gpurun_out/${R}_final_b64.json 2>> gpurun_out/${R}_final.err && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_b64.csv $CMD > gpurun_out/ncu_a.log 2>&1
# one chunk of 64 ciphertexts = 8 + 6 64-bit passes, ks_pass1 (digits -> auxiliary primes), the 32-bit passes (ntt_pass_kernel<.., IO32>: forward
# pass 2, 2 x inverse pass 2 + 1), aux_mac, 2 x aux_crt; the capture starts at the key upload and covers the warm-up step and the timed step
CMD2="python bench.py --steps 1 --warmup 1 --batch 64 --e2e-batch 2 --no-cpu-baseline --no-prof --no-ntt --no-chain"
timeout 200 $CMD2 > gpurun_out/${R}_final_b64_noprof.json 2>> gpurun_out/${R}_final.err && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'aux_|ks_pass1|ntt_pass' -c 56 -o /tmp/prof_${R}_aux -f $CMD2 > gpurun_out/ncu_b.log 2>&1 && ncu -i /tmp/prof_${R}_aux.ncu-rep --page raw --csv > gpurun_out/${R}_prof_aux_raw.csv 2>/dev/null && python tools/ncu_summary.py gpurun_out/${R}_prof_aux_raw.csv gpurun_out/${R}_ncu_full_ks_kernels_aux_final.json
tail -2 gpurun_out/ncu_b.log; tail -3 gpurun_out/${R}_final.err
