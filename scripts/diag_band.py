"""Diagnostic: which RoIs / bins of the tall-RoI input differ from the oracle, per kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from wssdl_bus_b200 import ops, synthetic as syn

B, H, W, C = 3, 38, 50, 96
bottom = syn.feature_map(40, B, H, W, C)
rng = np.random.default_rng(41)
tall = []
for i in range(60):
    x1 = rng.integers(-200, 700)
    y1 = rng.integers(-2500, 500)
    tall.append([i % B, x1, y1, x1 + rng.integers(16, 900), y1 + rng.integers(700, 4000)])
top_only = [[b, 10, 0, 400, 150] for b in range(B)]
bot_only = [[b, 300, 400, 790, 599] for b in range(B)]
rois = np.concatenate([np.array(tall + top_only + bot_only, np.float32),
                       syn.rois_for_pool(42, 200, B), syn.adversarial_rois(B, W, H)])
for mode in ("cpu", "gpu"):
    for PH, PW in ((7, 7), (2, 3), (14, 14)):
        wt, wa = oracle.clib.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16., bin_mode=0 if mode == "cpu" else 1)
        for kern in ("band", "direct", "tiled"):
            os.environ["WSSDL_ROI_FWD_KERNEL"] = kern
            # poison the blocks the caching allocator will hand out for top / argmax
            import torch
            pz = [torch.full((rois.shape[0], PH, PW, C), float("nan"), device="cuda"),
                  torch.full((rois.shape[0], PH, PW, C), -7, device="cuda", dtype=torch.int32)]
            torch.cuda.synchronize()
            del pz
            t, a = ops.roi_pool_forward(bottom, rois, PH, PW, 1 / 16., bin_mode=mode)
            t, a = t.cpu().numpy(), a.cpu().numpy()
            bad = np.argwhere((a != wa) | ~((t == wt) | (np.isnan(t) & np.isnan(wt))))
            print(mode, PH, PW, kern, "mismatches:", len(bad))
            if len(bad):
                rs = sorted(set(bad[:, 0].tolist()))
                print("  rois:", rs[:20])
                for r in rs[:4]:
                    bb = bad[bad[:, 0] == r]
                    print("  roi", r, rois[r], "bins(ph,pw):", sorted(set(map(tuple, bb[:, 1:3].tolist())))[:12],
                          "chan range", bb[:, 3].min(), bb[:, 3].max())
                    n, ph, pw, c = bb[0]
                    print("    got", t[n, ph, pw, c], a[n, ph, pw, c], "want", wt[n, ph, pw, c], wa[n, ph, pw, c])
