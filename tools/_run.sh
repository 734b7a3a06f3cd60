timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err
python -c "
import json
d=json.load(open('gpurun_out/bench_v13.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value']),'lk us',round(d['roofline']['us_per_launch']), d['clocks'])
"
tail -2 gpurun_out/bench_v13.err
