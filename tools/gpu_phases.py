"""per-phase cycle breakdown (needs the -DDSP_PHASES build named by DSP_LP_LIB)"""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
lib = S.load_library()
def phases(reset=True):
    buf = (C.c_ulonglong * 16)(); lib.dsp_lp_phases(buf, 1 if reset else 0); return np.array(list(buf), float)
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
t = TP.wind_battery(24)
names_band = {7: "loop/exit", 0: "residual+d", 1: "assemble", 2: "factor", 3: "newton pred", 4: "steps+mua", 5: "newton corr", 6: "steps+update"}
names_stage = {15: "loop/exit", 8: "residuals+check", 9: "recips+blocks", 10: "rhs pred", 11: "factor+fwd", 12: "back pred", 13: "recover+steps pred", 14: "rhs+solve corr"}
for kern, names, N in ((S.KERNEL_STAGE, names_stage, 10000), (S.KERNEL_BAND, names_band, 10000)):
    sol = S.BatchLPSolver(t, kernel=kern)
    phases(); sol.solve_host(lmp[:N], rp); ph = phases()
    tot = ph.sum()
    print("kernel", kern, "total cycles (warp 0 of block 0)", tot)
    for k, nm in names.items(): print("   %-22s %6.1f %%" % (nm, 100 * ph[k] / tot))
for nm, tt, cpv in (("nuclear", TP.nuclear(48), SC.c3(5000)), ("fossil", TP.fossil_surrogate(168), SC.c4(600))):
    sol = S.BatchLPSolver(tt); phases(); sol.solve_host(cpv, None); ph = phases(); tot = ph.sum()
    print(nm, "total", tot); [print("   %-22s %6.1f %%" % (n2, 100 * ph[k] / tot)) for k, n2 in names_band.items()]
