python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-single-thread --no-chain > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/r02q_bench.json
python bench.py --config cfg3 --steps 5 --warmup 3 --no-single-thread > gpurun_out/r02q_cfg3.json 2> gpurun_out/r02q_cfg3.err; echo "cfg3 rc=$?"; cut -c1-120 gpurun_out/r02q_cfg3.json
