#!/bin/bash
# Build K2 with several compile-time switches on the GPU box and time each: tools/scale_check.py
# (config-3 shape, k = 2..32) and bench.py (config 2).  Output: gpurun_out/k2_variants.log
out=gpurun_out/k2_variants.log
: > $out
for v in "" "-DDD_K2_MULHI=1" "-DDD_K2_VOTE_TAIL=1" "-DDD_K2_MULHI=1 -DDD_K2_VOTE_TAIL=1" "-DDD_SKETCH_THREADS_PLAIN=512" "-DDD_SKETCH_THREADS_SMALLK=512"; do
  echo "=== $v" >> $out
  DD_NVCC_EXTRA="$v" python -m dandd_b200.build --force >> $out 2>&1
  python tools/scale_check.py --sizes ${SIZES:-1e9} > /dev/null 2>&1   # lazy module loading out of the way
  python tools/scale_check.py --sizes ${SIZES:-1e9} 2>&1 | grep -o '"bases": [0-9]*\|"sketch_floor_ms": [0-9.]*\|"gbp_s_floor": [0-9.]*\|"oracle_k21_equal": [a-z]*\|"floor_equals_plain": [a-z]*' | paste -sd' ' >> $out
  python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench ms_per_step',d['ms_per_step'],'k2_ms',d['roofline']['launch_ms'],'e2e_ms',d['e2e']['ms_per_step'])" >> $out
done
python -m dandd_b200.build --force >> $out 2>&1
cat $out
