import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import knockoffgwas_b200 as kg
from knockoffgwas_b200 import synth
N, p, K, pairs = 21876, 60000, 100, 1094
prob = synth.make_problem_fast(N, p, K, seed=2022)
synth.add_ibd_pairs(prob, pairs)
related = np.concatenate([m for m, _ in prob.families])
unrelated = np.setdiff1d(np.arange(N, dtype=np.int32), related).astype(np.int32)
for name, kw in (("mixed", dict(unrelated=unrelated, ref_global=prob.ref, families=prob.families)),
                 ("unrel", dict(unrelated=unrelated)),
                 ("rel", dict(unrelated=np.zeros(0, np.int32), ref_global=prob.ref, families=prob.families))):
    for wide in ("0", "1"):
        os.environ["KGW_FORCE_WIDE"] = wide
        pl = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, **kw)
        pl.run(); pl.sync()
        ms = pl.run_timed(1)
        print(name, "wide=" + wide, round(ms, 1), pl.info(), flush=True)
        pl.close()
        kg.release_workspace()
