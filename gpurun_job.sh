for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -2 gpurun_out/bench_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); k=d['kernel_ms_per_step']; print($n, round(d['value'],2), round(d['ms_per_step'],3), round(d['e2e']['value'],2), {a: round(b,2) for a,b in k.items() if b>0.05}, d['lm_trace'][:2])"
done
