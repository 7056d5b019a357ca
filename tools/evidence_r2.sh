#!/bin/bash
# One B200: the round's single-GPU evidence (run under gpurun; outputs under gpurun_out/, copied to profiles/ afterwards).
set -u
export G2N_BENCH_NO_BIND=${G2N_BENCH_NO_BIND:-}
python bench.py > gpurun_out/r2_bench_C2.json 2> gpurun_out/r2_bench_C2.err
python tools/show_bench.py gpurun_out/r2_bench_C2.json | head -3
for c in "C3 1.0 5" "C4 1.0 5" "C4d 1.0 5" "C5 0.05 5"; do set -- $c; python bench.py --config $1 --scale $2 --steps $3 --warmup 3 > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err; python tools/show_bench.py gpurun_out/r2_bench_$1.json | head -2; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_C2_reference.json 2> gpurun_out/r2_bench_C2_reference.err; cut -c1-400 gpurun_out/r2_bench_C2_reference.json
python tools/bench_gz.py C2 > gpurun_out/r2_gz.txt 2>&1; tail -3 gpurun_out/r2_gz.txt
python tools/bench_ingest.py C2 > gpurun_out/r2_ingest.txt 2>&1; tail -1 gpurun_out/r2_ingest.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"k_tokenize|k_edges_scatter_flat|k_rows_sort|k_rows_write|k_assign_ids|k_mark_first" -s 24 -c 6 -o gpurun_out/r2_full_c2 python tools/kbench.py C2 1.0 2 > gpurun_out/r2_full_c2.log 2>&1; tail -1 gpurun_out/r2_full_c2.log | cut -c1-120
ncu --set full --clock-control none --import-source on -k regex:"k_tokenize" -s 3 -c 1 -o gpurun_out/r2_full_c5 python tools/kbench.py C5 0.05 2 > gpurun_out/r2_full_c5.log 2>&1; tail -1 gpurun_out/r2_full_c5.log | cut -c1-120
ncu --set full --clock-control none --import-source on -k regex:"k_tokenize" -s 3 -c 1 -o gpurun_out/r2_full_c3 python tools/kbench.py C3 1.0 2 > gpurun_out/r2_full_c3.log 2>&1; tail -1 gpurun_out/r2_full_c3.log | cut -c1-120
