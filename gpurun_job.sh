python -m pytest tests -m gpu -q 2>&1 | tail -4
for v in "" _B _C _D _E; do
  SAGE_BA_LIB=sage-slam_b200/lib/libsage_ba$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_var$v.json 2>gpurun_out/bench_var$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_var$v.json')); k=d['kernel_ms_per_step']; print('$v', round(d['value'],2), {a: round(b,2) for a,b in k.items() if b>0.2})"
done
