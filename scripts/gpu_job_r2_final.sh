set -x
python bench.py > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.err
python scripts/bench_kernels.py > /dev/null 2>&1 || true
NCU="ncu --set full --clock-control none --import-source on"
python scripts/prof_knn.py 4096 12 5 100 > gpurun_out/prof_knn_plain.log 2>&1 && $NCU -k regex:tile_kernel -s 2 -c 1 -f -o gpurun_out/r2_knn_set python scripts/prof_knn.py 4096 12 5 100 > gpurun_out/prof_knn_ncu.log 2>&1
python scripts/prof_large.py radius > gpurun_out/prof_large_radius.log 2>&1 && $NCU -k regex:gatq_large_x -s 2 -c 1 -f -o gpurun_out/r2_large_x_radius python scripts/prof_large.py radius > /dev/null 2>&1
$NCU -k regex:sim_step_grid -s 2 -c 1 -f -o gpurun_out/r2_sim_step_grid python scripts/prof_large.py radius > /dev/null 2>&1
python scripts/prof_large.py complete > gpurun_out/prof_large_complete.log 2>&1 && $NCU -k regex:gatq_large_x -s 2 -c 1 -f -o gpurun_out/r2_large_x_complete python scripts/prof_large.py complete > /dev/null 2>&1
python scripts/prof_stack.py > gpurun_out/prof_stack.log 2>&1 && $NCU -k regex:gatstack -s 5 -c 1 -f -o gpurun_out/r2_gatstack python scripts/prof_stack.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/prof_large_radius.log gpurun_out/prof_large_complete.log gpurun_out/prof_stack.log gpurun_out/prof_knn_plain.log
ls -la gpurun_out/*.ncu-rep | tail
