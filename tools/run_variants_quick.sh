# round-2 opener: `python tools/build_variants.py` here, then on the GPU box `bash tools/run_variants_quick.sh`
# (C2 timing + parity of every build/variants/*.so; one line each in gpurun_out/variants_quick.log)
for v in build/variants/*.so; do DSP_LP_LIB=/root/repo/$v timeout 120 python tools/gpu_quick_c2.py 2>&1 | grep -v Warn | tail -1; done | tee gpurun_out/variants_quick.log
