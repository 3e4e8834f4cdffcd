#!/bin/bash
# ab_bench.sh TAG variant...  -> runs bench.py (edge kernel only) for the default lib and each build/libporrt_<variant>.so
tag=$1; shift
python bench.py --steps 50 --warmup 3 --no-extras > gpurun_out/ab_${tag}_default.json 2> gpurun_out/ab_${tag}_default.err
python -c "import json;d=json.load(open('gpurun_out/ab_${tag}_default.json'));print('default', d['ms_per_step'], d['roofline']['frac'])"
for v in "$@"; do
  PORRT_B200_LIB=$PWD/build/libporrt_$v.so python bench.py --steps 50 --warmup 3 --no-extras > gpurun_out/ab_${tag}_$v.json 2> gpurun_out/ab_${tag}_$v.err
  python -c "import json;d=json.load(open('gpurun_out/ab_${tag}_$v.json'));print('$v', d['ms_per_step'], d['roofline']['frac'])" || tail -3 gpurun_out/ab_${tag}_$v.err
done
