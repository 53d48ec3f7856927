#!/bin/bash
# Multi-GPU evidence on ONE box (run under `gpurun --gpus 8`): the driver-visible multi-GPU tests, BASELINE configs[2]
# (rotations) at 2 / 4 / 8 GPUs, the contract's bench line at 8 GPUs (batch-sharded, with the per-rank e2e, the host copy
# probe and the single-process ckks_comm_* leg) and the optional limb-sharded mode at 8 GPUs.
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l); echo "gpus=$G"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_limb_shard.py -m gpu -x -q -k "every_gpu" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest.log
for N in 2 4 8; do
  [ $N -le $G ] || continue
  timeout 300 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --config cfg3 --steps 5 --warmup 3 --no-ntt > gpurun_out/r02m_cfg3_$N.json 2> gpurun_out/r02m_cfg3_$N.err; echo "cfg3 N=$N rc=$?"
done
for N in 8; do
  [ $N -le $G ] || continue
  timeout 400 $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02m_cfg4_$N.json 2> gpurun_out/r02m_cfg4_$N.err; echo "cfg4 N=$N rc=$?"; tail -c 300 gpurun_out/r02m_cfg4_$N.err
done
timeout 300 $TR --nproc-per-node $G --master-port 29800 bench.py --gpus $G --limb-sharded --batch 256 --ls-chunk 32 --steps 5 --warmup 3 > gpurun_out/r02m_limb_sharded_$G.json 2> gpurun_out/r02m_limb_sharded_$G.err; echo "lshard rc=$?"
