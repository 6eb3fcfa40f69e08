# on the GPU box: one C2 / C5-slice timing + parity line per build/variants/*.so
python tools/gpu_stage2_quick.py 2>&1 | tail -1 | tee gpurun_out/variants_quick.log
for v in build/variants/*.so; do DSP_LP_LIB=/root/repo/$v timeout 120 python tools/gpu_stage2_quick.py 2>&1 | grep -v Warn | tail -1; done | tee -a gpurun_out/variants_quick.log
