#!/bin/bash
# conv1 stage ablation: bit0 = no TMA, bit1 = no MMA, bit2 = no epilogue loads/stores
for d in 0 1 2 4 6 7; do
  I2L_CONV1_DBG=$d python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); print('dbg', $d, d['kernels_ms_per_step']['cnn.conv1_bf16'])"
done
