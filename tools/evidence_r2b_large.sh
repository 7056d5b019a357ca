#!/bin/bash
# One B200: the large configurations and the checked build again, on the final state of the partitioned row build.
set -u
R=${R:-r2b}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; tail -1 gpurun_out/${R}_pytest.log
for c in "C3 1.0 5" "C4d 1.0 5" "C5 0.05 5"; do set -- $c; python bench.py --config $1 --scale $2 --steps $3 --warmup 3 > gpurun_out/${R}_bench_$1.json 2> gpurun_out/${R}_bench_$1.err; python tools/show_bench.py gpurun_out/${R}_bench_$1.json 2>/dev/null | head -2; done
G2N_DBG_NOBUCKET=1 python tools/kbench.py C3 1.0 3 > gpurun_out/${R}_k_c3_rowrange.json 2>&1; G2N_DBG_NOBUCKET=1 python tools/kbench.py C4d 1.0 3 > gpurun_out/${R}_k_c4d_rowrange.json 2>&1
python tools/kbench.py C3 1.0 2 > gpurun_out/${R}_k_c3.json 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_tokenize|k_rows_count|k_bucket|k_sub_|k_rows_sort|k_rows_write|k_assign_ids" -s 33 -c 11 -f -o gpurun_out/${R}_full_c3 python tools/kbench.py C3 1.0 2 > gpurun_out/${R}_full_c3.log 2>&1; tail -1 gpurun_out/${R}_full_c3.log | cut -c1-120
G2N_LIB=build/var/libg2n_checked.so python tools/sanitize_run.py > gpurun_out/${R}_checked_run.log 2>&1; tail -1 gpurun_out/${R}_checked_run.log
G2N_LIB=build/var/libg2n_checked.so python -m pytest tests -m gpu -x -q > gpurun_out/${R}_checked_pytest.log 2>&1; tail -1 gpurun_out/${R}_checked_pytest.log
