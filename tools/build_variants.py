"""builds stage-kernel variants (compile-time switches) into build/variants/<name>.so for tools/run_variants.sh"""
import subprocess, sys, concurrent.futures as cf
sys.path.insert(0, ".")
from dispatches_b200.csrc import build as B
VARIANTS = {
    "v0_base": [],
    "v1_defer": ["-DDSP_DEFER_RCP"],
    "v2_base_r144": ["-DDSP_STAGE_WPB=2", "-DDSP_STAGE_MAXNREG=144"],
    "v3_defer_r144": ["-DDSP_DEFER_RCP", "-DDSP_STAGE_WPB=2", "-DDSP_STAGE_MAXNREG=144"],
    "v7_defer_r152": ["-DDSP_DEFER_RCP", "-DDSP_STAGE_WPB=1", "-DDSP_STAGE_MAXNREG=152"],
    "v4_base_ku4": ["-DDSP_STAGE_KU=4"],
    "v5_defer_ku4": ["-DDSP_DEFER_RCP", "-DDSP_STAGE_KU=4"],
    "v6_defer_ku1": ["-DDSP_DEFER_RCP", "-DDSP_STAGE_KU=1"],
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
        if "dsp_ipm_stage_wb_kernel" in l and "Compiling" in l:
            info = " | ".join(x.strip() for x in lines[i + 1:i + 4])
    return name, r.returncode, info
with cf.ThreadPoolExecutor(4) as ex:
    for name, rc, info in ex.map(one, VARIANTS.items()):
        print(name, rc, info)
