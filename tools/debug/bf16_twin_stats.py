"""rel-L2 / cosine of every gradient of the bf16 step replay vs the reference's fp32 step, with
the EMA normaliser in the contraction epilogue (late) and folded into the weights (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pytest, torch
import dusty_gan_v2_b200.functional as DF
import test_gpu_models as TM
from small_cfgs import D_MID, G_MID, sample_flat

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "trainer_step_mid.npz"), allow_pickle=True))
for late in (True, False):
    DF.set_late_ema(late)
    mp = pytest.MonkeyPatch()
    try:
        tr, rp = TM._replay_trainer(g, G_MID, D_MID, (32, 128), "bf16", False, mp)
        rp.run(0)
    finally:
        mp.undo()
    rows = []
    for grads, prefix in ((rp.g_grads, "gG_"), (rp.d_grads[0], "gD_"), (rp.d_grads[1], "gR1_")):
        for k, gr in grads.items():
            if prefix + k not in g:
                continue
            ref = torch.from_numpy(np.asarray(g[prefix + k])).float().reshape(-1)
            if float(ref.abs().max()) < 1e-9:
                continue
            got = sample_flat(gr.float().cpu()).reshape(-1)
            rows.append((float((got - ref).norm() / ref.norm()),
                         float(torch.dot(got, ref) / (got.norm() * ref.norm())), prefix + k))
    rows.sort(reverse=True)
    print("late_ema", late, "worst 6 of", len(rows))
    for r in rows[:6]:
        print("   rel %.4f cos %.5f %s" % r)
    for pfx in ("gG_", "gD_", "gR1_"):
        v = [r[0] for r in rows if r[2].startswith(pfx)]
        print("   ", pfx, "median rel %.4f  max %.4f" % (float(np.median(v)), max(v)))
