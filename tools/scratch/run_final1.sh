set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1 || true
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_partition_scatter_bulk|k_probe_inner_u1_seg" --launch-skip 6 -c 2 -o gpurun_out/r1_pipeline -f python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/r1_pipeline.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches_pipeline.csv python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/r1_launches_pipeline.out 2>&1
timeout 400 python bench.py > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err
