python tools/gpu_band_rw_check.py 2>&1 | tail -6 | tee gpurun_out/band_rw.log
DSP_BAND_NO_RW=1 python tools/gpu_band_rw_check.py 2>&1 | tail -6 | tee -a gpurun_out/band_rw.log
python -m pytest tests/test_solar_battery_hydrogen.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_solar.log
