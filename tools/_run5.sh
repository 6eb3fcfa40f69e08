python -m pytest tests/test_solar_battery_hydrogen.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_solar.log
DSP_LP_LIB=build/variants/libdsp_phases.so python tools/gpu_phases.py 2>&1 | tail -34 > gpurun_out/phases_ws.log
