"""builds compile-time variants of the library into build/variants/libdsp_<name>.so (experiments; run them with DSP_LP_LIB=...)"""
import subprocess, sys, concurrent.futures as cf
sys.path.insert(0, ".")
from dispatches_b200.csrc import build as B
VARIANTS = {
    "s2_base": [],
    "s2_rcp1": ["-DDSP_S2_RCP_NEWTON=1"],             # one Newton step after the reciprocal seed instead of two
    "s2_warps12": ["-DDSP_S2_WARPS=12"],              # 170-register cap; run with DSP_STAGE2_GEOM=16,2 (two periods per lane)
    "s2_warps10": ["-DDSP_S2_WARPS=10"],              # 204-register cap
    "s2_sync31": ["-DDSP_S2_SYNCMASK=31"],
}
out = B.ROOT / "build" / "variants"
out.mkdir(parents=True, exist_ok=True)
for f in out.glob("*.so"):
    f.unlink()
def one(item):
    name, flags = item
    cmd = [B.nvcc_path(), *B.NVCC_FLAGS, "-Xptxas", "-v", *flags, *map(str, B.SOURCES), "-o", str(out / f"libdsp_{name}.so")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    lines = r.stderr.splitlines()
    info = ""
    for i, l in enumerate(lines):
        if "dsp_ipm_stage2_wb_kernelILi8ELi3ELb1" in l and "Compiling" in l:
            info = " | ".join(x.strip() for x in lines[i + 1:i + 4])
    return name, r.returncode, info
with cf.ThreadPoolExecutor(4) as ex:
    for name, rc, info in ex.map(one, VARIANTS.items()):
        print(name, rc, info)
