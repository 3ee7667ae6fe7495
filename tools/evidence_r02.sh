#!/bin/bash
# Round-2 evidence on ONE B200 (run under gpurun): launch lists of the bench commands and ncu --set full captures of the final
# kernels.  Every program is first run without ncu; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch lists (gpu__time_duration per launch) of the T* and configs[4] bench commands
python bench.py --steps 1 --warmup 1 --no-cpu > $O/ev_tstar_plain.json 2> $O/ev_tstar_plain.err && \
  $NCU --metrics gpu__time_duration.sum -c 2000 --csv --log-file $O/ev_launches_tstar.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/ev_tstar_ncu.json 2> $O/ev_tstar_ncu.err
python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu > $O/ev_c5_plain.json 2> $O/ev_c5_plain.err && \
  $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/ev_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu > $O/ev_c5_ncu.json 2> $O/ev_c5_ncu.err
# 2. --set full captures of the final kernels (each program exits 0 without ncu first)
python tools/bench_schur.py 1000064 1000,200 > $O/ev_schur_plain.json && \
  $NCU --set full --import-source on -k regex:gemm_kernel --launch-skip 3 -c 1 -f -o $O/ev_schur_f200 python tools/bench_schur.py 1000064 1000,200 > /dev/null 2>&1
python tools/bench_schur.py 125056 1000,200 > $O/ev_schur8_plain.json && \
  $NCU --set full --import-source on -k regex:gemm_kernel --launch-skip 6 -c 2 -f -o $O/ev_schur_tail python tools/bench_schur.py 125056 1000,200 > /dev/null 2>&1
python tools/bench_eval.py gaussian 1000000 50 200 > $O/ev_evalp_plain.json && \
  $NCU --set full --import-source on -k regex:eval_panel_kernel --launch-skip 1 -c 1 -f -o $O/ev_evalp_f200 python tools/bench_eval.py gaussian 1000000 50 200 > /dev/null 2>&1
python tools/bench_eval.py laplace 1000000 50 200 > $O/ev_direct_plain.json && \
  $NCU --set full --import-source on -k regex:direct_rows_tma --launch-skip 1 -c 1 -f -o $O/ev_direct_l1 python tools/bench_eval.py laplace 1000000 50 200 > /dev/null 2>&1
python tools/bench_syrk.py 1000 100000 > $O/ev_syrk_plain.json && \
  $NCU --set full --import-source on -k regex:syrk_kernel --launch-skip 1 -c 1 -f -o $O/ev_syrk python tools/bench_syrk.py 1000 100000 > /dev/null 2>&1
ls -la $O/ev_*
