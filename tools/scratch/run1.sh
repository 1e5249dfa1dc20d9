set -x
timeout 600 python -m pytest tests/test_gpu_join.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_join.txt
timeout 200 python bench.py --skip-e2e --skip-cpu > gpurun_out/bench_part1.json 2> gpurun_out/bench_part1.err
TG_PROBE_SEG_VEC=0 timeout 200 python bench.py --skip-e2e --skip-cpu > gpurun_out/bench_part1_sv0.json 2> gpurun_out/bench_part1_sv0.err
TG_PROBE_PARTS=10 timeout 200 python bench.py --skip-e2e --skip-cpu > gpurun_out/bench_part1_p10.json 2> gpurun_out/bench_part1_p10.err
TG_PROBE_PARTS=14 timeout 200 python bench.py --skip-e2e --skip-cpu > gpurun_out/bench_part1_p14.json 2> gpurun_out/bench_part1_p14.err
TG_PROBE_PARTS=16 timeout 200 python bench.py --skip-e2e --skip-cpu > gpurun_out/bench_part1_p16.json 2> gpurun_out/bench_part1_p16.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/part_launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/part_launches.out 2>&1
