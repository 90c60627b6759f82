timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1c_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r1c_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
