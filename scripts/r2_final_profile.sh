# final round-2 evidence: launch lists (bench step, PRM build) + full captures of the edge kernel, the one-pass
# thread-per-query radius kernel of the PRM build and the radius tile kernel.  Each program first runs without ncu.
set -x
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/b_noextras.json 2> gpurun_out/b_noextras.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edge_validity_v3 -c 1 --launch-skip 2 -f -o gpurun_out/r2_edge python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_edge.log 2>&1
python scripts/prm_only.py 1000000 2>&1 | tail -2 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_prm_launches.csv python scripts/prm_only.py 1000000 > gpurun_out/ncu_prm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"radius_kernel|radius_wide_kernel" -c 2 --launch-skip 2 -f -o gpurun_out/r2_prm_radius python scripts/prm_only.py 1000000 > gpurun_out/ncu_prm_radius.log 2>&1
python scripts/nn_radius_run.py 2 2>&1 | tail -2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:nt_radius -c 2 -f -o gpurun_out/r2_radius python scripts/nn_radius_run.py 1 > gpurun_out/ncu_radius.log 2>&1
ls -la gpurun_out/*.ncu-rep
