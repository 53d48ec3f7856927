timeout 120 tools/tc_ntt_bench > gpurun_out/r02e_tc_ntt.json 2> gpurun_out/r02e_tc_ntt.err; echo "tc rc=$?"; cat gpurun_out/r02e_tc_ntt.json | head -c 1800; tail -c 300 gpurun_out/r02e_tc_ntt.err
python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log; tail -4 gpurun_out/r02e_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02e_bench.err
python bench.py --steps 4 --warmup 2 --no-ntt --no-chain --no-single-thread --e2e-batch 14 --ks-scratch-mib 8192 > gpurun_out/r02e_scr8g.json 2> gpurun_out/r02e_scr8g.err; echo "scr8g rc=$?"
python bench.py --steps 4 --warmup 2 --no-ntt --no-chain --no-single-thread --e2e-batch 14 --ks-scratch-mib 2048 > gpurun_out/r02e_scr2g.json 2> gpurun_out/r02e_scr2g.err; echo "scr2g rc=$?"
python bench.py --config cfg3 --steps 5 --warmup 3 --no-single-thread > gpurun_out/r02e_cfg3.json 2> gpurun_out/r02e_cfg3.err; echo "cfg3 rc=$?"; tail -c 400 gpurun_out/r02e_cfg3.err
python bench.py --ntt-sweep --steps 3 --warmup 1 > gpurun_out/r02e_ntt_sweep.json 2> gpurun_out/r02e_ntt_sweep.err; echo "sweep rc=$?"
