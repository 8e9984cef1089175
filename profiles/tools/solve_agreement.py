"""Agreement of the production kernels with the oracle (iteration counts, poses) per solve variant."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, oracle_run, gpu_run, check_parity
from oracle import oracle as O
orc = O.OracleLib()
cases = {"c1 psz8 P100": dict(seed=31, ntracks=64), "c1 donorm": dict(seed=32, ntracks=64, donorm=1),
         "c1 donorm+patchnorm": dict(seed=33, ntracks=64, donorm=1, dopatchnorm=1),
         "c3 psz32 P4": dict(seed=21, w=1920, h=1080, psz=32, npts=4, ntracks=256)}
for name, kw in cases.items():
    case = make_case(**kw)
    o = oracle_run(orc, case, trace_cap=48)
    o2 = oracle_run(orc, case, trace_cap=48, sum_mode=1)
    g = gpu_run(ict, case, trace_cap=48)
    m = check_parity(g, o, case, gates=False); s = check_parity(o2, o, case, gates=False)
    dp = np.abs(g["p_out"] - o["p_out"]).max(axis=1); ds = np.abs(o2["p_out"] - o["p_out"]).max(axis=1)
    print("%-22s gpu: same %.3f jtr1 %.1e med_tr %.1e rot %.1e pose med %.1e p99 %.1e | oracle AVX-vs-SSE: same %.3f med_tr %.1e pose med %.1e p99 %.1e"
          % (name, m["frac_same"], m["jtr_first"], m["median_tr"], m["worst_rot"], np.median(dp), np.percentile(dp, 99),
             s["frac_same"], s["median_tr"], np.median(ds), np.percentile(ds, 99)))
