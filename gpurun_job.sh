python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_refcpu.json 2> gpurun_out/bench_r1_refcpu.err; tail -2 gpurun_out/bench_r1_refcpu.err
python bench.py --impl reference-gpu --steps 4 --warmup 1 > gpurun_out/bench_r1_refgpu.json 2> gpurun_out/bench_r1_refgpu.err; tail -2 gpurun_out/bench_r1_refgpu.err
python bench.py --impl reference-gpu --steps 8 --warmup 2 --ref-samples 3072 > gpurun_out/bench_r1_refgpu_n3072.json 2> gpurun_out/bench_r1_refgpu_n3072.err; tail -2 gpurun_out/bench_r1_refgpu_n3072.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^photo_kernel -s 6 -c 2 -f -o gpurun_out/prof_photo_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_photo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^geo_kernel -s 6 -c 2 -f -o gpurun_out/prof_geo_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_geo.log 2>&1
for f in bench_r1_n1 bench_r1_refcpu bench_r1_refgpu bench_r1_refgpu_n3072; do echo $f; cut -c1-900 gpurun_out/$f.json; done
