#!/bin/bash
# Round-2 ncu captures (one GPU, --clock-control none); raw reports land in gpurun_out/, summaries are written by tools/summarize_ncu.py
set -x
NCU="ncu --set full --import-source on --clock-control none"
timeout 400 $NCU -k regex:"cb_select_count|cb_pipeline_select" --launch-skip 2 --launch-count 2 -o gpurun_out/r2_config1_f64_full -f python bench.py --workload config1 --variant f64 --steps 1 --warmup 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 $NCU -k regex:"cb_pipeline_agg|cb_finalize|k_partition_ids|k_pid_place|k_pid_block_hist|k_gather_rows|k_gather_bits" --launch-skip 16 --launch-count 16 -o gpurun_out/r2_groupby_full -f python bench.py --workload groupby --steps 1 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/ncu_b.log 2>&1
timeout 600 $NCU -k regex:"k_pq_" --launch-skip 20 --launch-count 14 -o gpurun_out/r2_parquet_full -f python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
