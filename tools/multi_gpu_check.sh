set -x
nvidia-smi topo -m > gpurun_out/r2p_topo.txt 2>&1
python tools/h2d_bw.py 1 2 4 8 > gpurun_out/r2p_h2d_bw.txt 2>&1
for n in 8 2; do
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2p_bench_${n}gpu.json 2> gpurun_out/r2p_bench_${n}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/r2p_bench_ref_${n}gpu.json 2>> gpurun_out/r2p_bench_${n}gpu.err
done
grep -c "NCCL INFO" gpurun_out/r2p_bench_8gpu.err; grep -i "comm.*nranks\|Init COMPLETE" gpurun_out/r2p_bench_8gpu.err | head -3
cat gpurun_out/r2p_h2d_bw.txt
python - <<'P'
import json
for n in (8,2):
    try:
        d=json.loads(open(f'gpurun_out/r2p_bench_{n}gpu.json').read().strip().splitlines()[-1])
        print(n, d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('parity'))
        r=json.loads(open(f'gpurun_out/r2p_bench_ref_{n}gpu.json').read().strip().splitlines()[-1])
        print(' ref', r['value'], r['cpu_baseline']['cores'])
    except Exception as e: print(n, 'ERR', e)
P
