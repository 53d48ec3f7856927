#!/bin/bash
# e2e throughput of the host-buffer entry point vs pipeline chunk size (cfg4, 128 ciphertext pairs per call)
for c in 32 64 128 256; do
  python bench.py --steps 3 --warmup 1 --batch 28 --e2e-batch 128 --no-cpu-baseline --no-prof --host-chunk-mib $c 2>&1 | tail -1 > /tmp/e2e_$c.json
  python -c "
import json,sys; d=json.load(open('/tmp/e2e_$c.json')); print('chunk_mib', $c, 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2))"
done
