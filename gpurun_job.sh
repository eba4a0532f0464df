python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; tail -15 gpurun_out/gpu_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -3 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:photo_kernel -s 2 -c 2 -f -o gpurun_out/prof_photo python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_photo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:geo_kernel -s 2 -c 2 -f -o gpurun_out/prof_geo python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_geo.log 2>&1
ls -la gpurun_out
