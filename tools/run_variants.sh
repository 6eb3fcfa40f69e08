for v in build/variants/*.so; do DSP_LP_LIB=/root/repo/$v timeout 150 python tools/gpu_variant_bench.py 2>&1 | grep -v Warn | tail -1; done | tee gpurun_out/variants_r2.log
