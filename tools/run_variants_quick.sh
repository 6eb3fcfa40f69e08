# on the GPU box: one C2 / C5-slice timing + parity line per build/variants/*.so
rm -f gpurun_out/variants_quick.log
for v in build/variants/*.so; do DSP_LP_LIB=/root/repo/$v timeout 120 python tools/gpu_stage2_quick.py 2>&1 | grep -v Warn | tail -1; done | tee -a gpurun_out/variants_quick.log
for w in 12 10; do DSP_STAGE2_GEOM=16,2 DSP_LP_LIB=/root/repo/build/variants/libdsp_s2_warps$w.so timeout 120 python tools/gpu_stage2_quick.py 2>&1 | grep -v Warn | tail -1 | sed "s/^/geom 16,2 /"; done | tee -a gpurun_out/variants_quick.log
DSP_STAGE2_GEOM=16,2 DSP_LP_LIB=/root/repo/build/variants/libdsp_s2_base.so timeout 120 python tools/gpu_stage2_quick.py 2>&1 | grep -v Warn | tail -1 | sed "s/^/geom 16,2 /" | tee -a gpurun_out/variants_quick.log
