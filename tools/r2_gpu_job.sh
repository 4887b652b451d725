set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_fulltests2.log 2>&1; tail -5 gpurun_out/r2_fulltests2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 400 gpurun_out/r2_bench_default.json
for c in 64 96; do python bench.py --steps 4 --warmup 3 --chunk $c --no-extra --no-cpu-baseline > gpurun_out/r2_bench_chunk$c.json 2>&1; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_chunk$c.json').read().strip().splitlines()[-1]); print('chunk $c', d['value'], d['e2e']['value'], d['clocks'])"; done
python bench.py --steps 5 --warmup 3 --precision mixed --no-extra --no-cpu-baseline > gpurun_out/r2_bench_mixed_1h.json 2>&1; tail -c 300 gpurun_out/r2_bench_mixed_1h.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_x3.csv python bench.py --hours 0.02 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_ncu_x3.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_mixed.csv python bench.py --hours 0.02 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --precision mixed > gpurun_out/r2_ncu_mixed.log 2>&1
ls -la gpurun_out | tail -12
