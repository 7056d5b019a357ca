#!/bin/bash
# One B200: the round's single-GPU evidence (run under gpurun; outputs under gpurun_out/, summarised under profiles/ afterwards).
set -u
R=${R:-r2b}
python bench.py > gpurun_out/${R}_bench_C2.json 2> gpurun_out/${R}_bench_C2.err
python tools/show_bench.py gpurun_out/${R}_bench_C2.json | head -3
for c in "C3 1.0 5" "C4 1.0 5" "C4d 1.0 5" "C5 0.05 5"; do set -- $c; python bench.py --config $1 --scale $2 --steps $3 --warmup 3 > gpurun_out/${R}_bench_$1.json 2> gpurun_out/${R}_bench_$1.err; python tools/show_bench.py gpurun_out/${R}_bench_$1.json | head -2; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_C2_reference.json 2> gpurun_out/${R}_bench_C2_reference.err; cut -c1-400 gpurun_out/${R}_bench_C2_reference.json
python bench.py --steps 2 --warmup 3 > gpurun_out/${R}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${R}_ncu_launches.log 2>&1
tail -2 gpurun_out/${R}_ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"k_tokenize|k_edges_scatter_flat|k_rows_sort|k_rows_write|k_assign_ids|k_mark_first" -s 24 -c 6 -o gpurun_out/${R}_full_c2 python tools/kbench.py C2 1.0 2 > gpurun_out/${R}_full_c2.log 2>&1; tail -1 gpurun_out/${R}_full_c2.log | cut -c1-120
python tools/kbench.py C3 1.0 2 > gpurun_out/${R}_k_c3.json 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_tokenize|k_rows_count|k_bucket|k_sub_|k_rows_sort|k_rows_write|k_assign_ids" -s 33 -c 11 -o gpurun_out/${R}_full_c3 python tools/kbench.py C3 1.0 2 > gpurun_out/${R}_full_c3.log 2>&1; tail -1 gpurun_out/${R}_full_c3.log | cut -c1-120
# bounds-checked build (compute-sanitizer is closed on this pool): the small fixtures and the whole GPU suite
G2N_LIB=build/var/libg2n_checked.so python tools/sanitize_run.py > gpurun_out/${R}_checked_run.log 2>&1; tail -1 gpurun_out/${R}_checked_run.log
G2N_LIB=build/var/libg2n_checked.so python -m pytest tests -m gpu -x -q > gpurun_out/${R}_checked_pytest.log 2>&1; tail -1 gpurun_out/${R}_checked_pytest.log
