import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine, SCHEDULE_WALKER, SCHEDULE_LEVELLED
from oracle.oracle_engine import OracleEngine
tree, model, pat, freq, inv = make_case("aa_c4", ntaxa=40, nsites=700, seed=11, mean_len=0.15)
for sched in (SCHEDULE_WALKER, SCHEDULE_LEVELLED):
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=True, schedule=sched)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=True)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    g.branch_logl(ops, a, b, L); o.branch_logl(ops, a, b, L)
    worst = (0, None)
    for k, op in enumerate(ops):
        s = int(op["dst_slot"])
        G, O = g.get_partial(s), o.get_partial(s)
        sc = o.get_scale_num(s)
        P = G.shape[0]
        G = G.reshape(P, model.ncat_mix, 20); O = O.reshape(P, model.ncat_mix, 20)
        scale = np.abs(O).reshape(P, -1).max(axis=1); scale[scale == 0] = 1
        err = np.abs(G - O).max(axis=2) / scale[:, None]
        i = np.unravel_index(err.argmax(), err.shape)
        if err.max() > worst[0]:
            worst = (err.max(), (k, s, i, sc.reshape(P, -1)[i[0]], np.abs(O[i[0]]).max(axis=1), G[i[0], i[1]], O[i[0], i[1]]))
    print("schedule", sched, "worst err", worst[0])
    k, s, i, scn, catmax, gv, ov = worst[1]
    print(" op", k, "slot", s, "pattern,cat", i, "scale_num row", scn, "\n per-cat max", catmax, "\n gpu", gv, "\n ora", ov)
    g.close()
