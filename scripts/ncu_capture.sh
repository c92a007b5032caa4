#!/bin/bash
# Measurements and ncu captures of round 2 (run on the GPU box through gpurun; one GPU).  Outputs under gpurun_out/,
# summarised into profiles/ by scripts/ncu_summary.py.  Every command has run (exit 0) WITHOUT ncu before.
#   make -C phylo_b200/csrc tools && gpurun --timeout 900 -- 'bash scripts/ncu_capture.sh [all]'
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# the bench line, the scoring kernels on controlled mixes of child pairs, the inner loop's ceiling, the event kernel's phases
python bench.py > $O/r2_bench_gtr.json 2> $O/r2_bench_gtr.err
scripts/score_bench 20 > $O/r2_score_bench.txt 2>&1
scripts/rows_micro > $O/r2_rows_micro.txt 2>&1
python scripts/event_phases.py 64 10000 65536 0 > $O/r2_event_phases_64k.txt 2>&1
# launch list of the bench command (per-launch times: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-eager-dense --no-cpu-baseline --no-hbm-kernels > $O/r2_ncu_bench.log 2>&1
# the scoring kernels and the event kernel: a late (heavy) rank event of the second 64 x 10k x 65,536 forward
$NCU -k regex:merge_score_rows -s 113 -c 1 -o $O/r2_full_rows python scripts/fwd_only.py 64 10000 65536 0 > $O/r2_full_rows.log 2>&1
$NCU -k regex:merge_score_kernel -s 118 -c 1 -o $O/r2_full_generic python scripts/fwd_only.py 64 10000 65536 0 > $O/r2_full_generic.log 2>&1
$NCU -k regex:lz_event_kernel -s 100 -c 1 -o $O/r2_full_event python scripts/fwd_only.py 64 10000 65536 0 > $O/r2_full_event.log 2>&1
if [ "${1:-}" = "all" ]; then
  # the HBM-bound kernels on distinct children (scripts/hbm_kernels.py), the sparse reverse pass, the look-ahead kernel
  $NCU -k regex:merge_fwd_kernel -s 2 -c 1 -o $O/r2_full_hbm_merge_fwd python scripts/hbm_kernels.py > $O/r2_full_hbm_fwd.log 2>&1
  $NCU -k regex:merge_bwd_kernel -s 2 -c 1 -o $O/r2_full_hbm_merge_bwd python scripts/hbm_kernels.py > $O/r2_full_hbm_bwd.log 2>&1
  $NCU -k regex:bwd_sparse_kernel -s 1 -c 1 -o $O/r2_full_bwd_sparse python scripts/fwd_only.py 64 10000 65536 0 bwd > $O/r2_full_bwd_sparse.log 2>&1
  $NCU -k regex:lookahead_kernel -s 20 -c 1 -o $O/r2_full_lookahead python scripts/perf_probe.py 17 3260 4096 0 skip 0 10 > $O/r2_full_lookahead.log 2>&1
fi
ls -la $O/r2_full_*.ncu-rep
