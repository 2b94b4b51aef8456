n=$1
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r3f_bench_${n}gpu.json 2> gpurun_out/r3f_bench_${n}gpu.err
echo rc=$?
python - <<P
import json
d=json.loads(open('gpurun_out/r3f_bench_${n}gpu.json').read().strip().splitlines()[-1])
print($n, round(d['value']), round(d['e2e']['value']), round(d['e2e']['pageable']['value']), round(d['roofline']['frac'],3), d.get('parity'))
P
grep -c "NCCL INFO" gpurun_out/r3f_bench_${n}gpu.err
