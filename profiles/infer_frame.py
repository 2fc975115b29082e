#!/usr/bin/env python
"""configs[2] under ncu: `frame` = 3 inference frames (248x184, device-driven rounds), `update` = 3 occupancy-grid updates.
usage: ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file X.csv python profiles/infer_frame.py frame"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from customnerf_b200 import trainer, synthetic as syn  # noqa: E402

dev = torch.device("cuda")
what = sys.argv[1] if len(sys.argv) > 1 else "frame"
model = trainer.build_scene_model(dev)
if what == "frame":
    model.eval()
    o, d = syn.camera_rays(184, 248)
    o, d = o.to(dev), d.to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for _ in range(3):
            out = model.render(o[None], d[None], perturb=False)
    torch.cuda.synchronize()
    print("frame ok", float(out["image"].float().mean()))
else:
    model.train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for _ in range(3):
            model.update_extra_state()
    torch.cuda.synchronize()
    print("update ok", model.mean_density)
