import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
for k in a.files:
    if not k.endswith("_v"): continue
    lv = k[:-2]
    sa, sb = a[lv + "_s"], b[lv + "_s"]
    both = (sa == 0) & (sb == 0)
    dv = np.abs(a[lv + "_v"][both] - b[lv + "_v"][both])
    qa, qb = a[lv + "_q"][both], b[lv + "_q"][both]
    same_nan = (np.isnan(qa) == np.isnan(qb)).all()
    fin = np.isfinite(qa) & np.isfinite(qb)
    dq = np.abs(qa[fin] - qb[fin])
    inf_mismatch = int((np.isinf(qa) != np.isinf(qb)).sum())
    print(lv, "both ok", int(both.sum()), "tree-only ok", int(((sa == 0) & (sb != 0)).sum()), "old-only ok", int(((sa != 0) & (sb == 0)).sum()),
          "max dV %.2g" % (dv.max() if dv.size else 0), "max dQ %.2g" % (dq.max() if dq.size else 0), "nan pattern same", bool(same_nan), "inf mismatches", inf_mismatch,
          "status differs (2 vs other)", int(((sa == 2) != (sb == 2)).sum()))
    if inf_mismatch:
        idx = np.argwhere(np.isinf(qa) != np.isinf(qb))[:5]
        for i in idx: print("   ", i, qa[i[0]][i[1]], qb[i[0]][i[1]])
