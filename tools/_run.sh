python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('N=2 value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'n_gpus',d['n_gpus'],d['scaling'], d['clocks'])
"
tail -3 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
