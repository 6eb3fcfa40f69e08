"""builds stage-kernel variants (compile-time switches) into build/variants/<name>.so for tools/run_variants.sh"""
import subprocess, sys, concurrent.futures as cf
sys.path.insert(0, ".")
from dispatches_b200.csrc import build as B
VARIANTS = {
    "r2_base": [],
    "r2_hybrid2": ["-DDSP_EXPERIMENT_HYBRID2"],      # run with DSP_BAND_MODE=hybrid2 (tools/gpu_band_modes.py)
    "r2_park_168": ["-DDSP_STAGE_PARK=1"],             # parked temporaries at the current 168-register / 12-warp point
    "r2_park_128": ["-DDSP_STAGE_PARK=1", "-DDSP_STAGE_MINB=4"],     # 128 registers / 16 warps per SM
    "r2_park2_128": ["-DDSP_STAGE_PARK=2", "-DDSP_STAGE_MINB=4"],    # 40 parked doubles per lane
    "r2_park2_96": ["-DDSP_STAGE_PARK=2", "-DDSP_STAGE_MINB=5"],     # 96 registers / 20 warps per SM (5 CTAs x 4 warps, 40 KB smem each)
    "r2_park2_112": ["-DDSP_STAGE_PARK=2", "-DDSP_STAGE_WPB=2", "-DDSP_STAGE_MINB=9"],   # ptxas picks 96 registers / 18 warps per SM
    "r2_nopark_128": ["-DDSP_STAGE_MINB=4"],
    "r2_start1": ["-DDSP_STAGE_START=1"],            # primal-feasible start; time it with BatchLPSolver(step_frac=0.99995) too
    # per-phase cycle counters (tools/gpu_phases.py with DSP_LP_LIB): where do 16 warps per SM lose what they gain?
    "r2_phases_base": ["-DDSP_PHASES"],
    "r2_phases_park_128": ["-DDSP_PHASES", "-DDSP_STAGE_PARK=1", "-DDSP_STAGE_MINB=4"],
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
