"""builds compile-time variants of the library into build/variants/libdsp_<name>.so (experiments; run them with DSP_LP_LIB=...)"""
import subprocess, sys, concurrent.futures as cf
sys.path.insert(0, ".")
from dispatches_b200.csrc import build as B
VARIANTS = {
    # generation-2 stage kernel: which phase boundaries of an IPM round carry a CTA barrier (bit k = boundary k of 5)
    "s2_sync31": [],
    "s2_sync21": ["-DDSP_S2_SYNCMASK=21"],
    "s2_sync5": ["-DDSP_S2_SYNCMASK=5"],
    "s2_sync1": ["-DDSP_S2_SYNCMASK=1"],
    "s2_sync0": ["-DDSP_S2_SYNCMASK=0"],           # only the exit vote at the top of the round
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
