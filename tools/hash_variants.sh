#!/bin/bash
# Tuning aid: time the hash-aggregate kernels of `bench.py --workload groupby` under JIT experiment defines / thread counts.
run() { # name defs threads
  CB200_JIT_DEFS="$2" CB200_HASH_THREADS="$3" timeout 200 python bench.py --workload groupby --rows ${ROWS:-250000000} --steps 3 --warmup 1 --no-e2e --no-cpu --no-check 2>/dev/null \
   | python -c "import sys,json; l=json.loads(sys.stdin.readline()); print('$1', 'step_ms=%.1f'%l['ms_per_step'], l['phases_ms'], 'partial_ms_per_launch=%.2f'%l['roofline']['ms_per_launch'])"
}
run base "" ""
run nocarry "CB_X_NOCARRY=1" ""
run gidblock32 "CB_X_GIDBLOCK=32" ""
run both "CB_X_NOCARRY=1;CB_X_GIDBLOCK=32" ""
run both960 "CB_X_NOCARRY=1;CB_X_GIDBLOCK=32" 960
run base960 "" 960
run base256 "" 256
