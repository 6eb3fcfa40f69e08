python -m pytest tests/test_solar_battery_hydrogen.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_solar.log
ncu --set full --clock-control none --import-source on -k regex:stage2_wb -s 1 -c 1 -o gpurun_out/prof_r2_stage2_d -f python tools/gpu_one_stage2.py > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 10 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
