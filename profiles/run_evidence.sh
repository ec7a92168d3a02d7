set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
# launch list of the bench command (per-launch durations, cold/serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 1 --batch 100 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
# full capture of the four per-iteration kernels
ncu --set full --clock-control none --import-source on -k regex:"k_fwd_data|k_adj_tile|k_update|k_fwd_sym" -s 8 -c 4 -o gpurun_out/prof_r1_final python profiles/prof_run.py 100 6 > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out | tail -4
# explicit-row path (trilinear, cfg2 shape): launch list of one candidate + full capture of its two operators
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_trilinear.csv python profiles/linear_breakdown.py 256 > gpurun_out/ncu_launch_tri.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fwd_csr|k_adj_csc|k_exp_rows" -c 6 -o gpurun_out/prof_r1_explicit -f python profiles/linear_breakdown.py 256 > gpurun_out/ncu_explicit.log 2>&1
