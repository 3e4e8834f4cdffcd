set -x
python -m pytest tests -m gpu -x -q -k "reachable or belief" 2>&1 | tail -2
python scripts/c4_profile.py 2>&1 | tail -2
python scripts/nn_radius_run.py 2 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:nt_radius -c 1 -f -o gpurun_out/r2_radius python scripts/nn_radius_run.py 1 > gpurun_out/ncu_radius.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:colsolve_push -c 12 -f -o gpurun_out/r2_colsolve python scripts/c4_profile.py > gpurun_out/ncu_colsolve.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/b_noextras.json 2> gpurun_out/b_noextras.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edge_validity_v3 -c 1 --launch-skip 2 -f -o gpurun_out/r2_edge python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_edge.log 2>&1
ls -la gpurun_out/*.ncu-rep
