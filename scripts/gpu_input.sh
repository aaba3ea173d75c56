#!/bin/bash
# Runs on the GPU box (under gpurun): input-pipeline parity tests, default bench, ncu capture of the augmentation kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_gpu.py -m gpu -x -q > gpurun_out/pytest_data_gpu.log 2>&1; echo "pytest data rc=$?" | tee -a gpurun_out/pytest_data_gpu.log
tail -15 gpurun_out/pytest_data_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
print('input', json.dumps(d.get('input'))[:1500])
"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:augment_kernel -s 2 -c 1 -f -o gpurun_out/r02_augment \
    python scripts/input_prof.py > gpurun_out/ncu_augment.log 2>&1; echo "ncu rc=$?"
timeout 120 ncu -i gpurun_out/r02_augment.ncu-rep --page raw --csv > gpurun_out/r02_augment_raw.csv 2>/dev/null
