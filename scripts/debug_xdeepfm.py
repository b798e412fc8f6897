"""GPU debug: per-parameter gradient errors of xdeepfm against the oracle, per CIN precision."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import make_golden as mg
from oracle import models as om
import test_gpu_models as T

cuda = torch.device("cuda", 0)
for prec in ("fp32", "tf32x3"):
    spec = mg.small_spec()
    p64 = om.init_params("xdeepfm", spec.total_rows, deep_layers=(32, 16), seed=3, cin_layers=(16, 8))
    feats, batch = mg.model_batch("xdeepfm", 64, 7, spec)
    out64, g64 = om.loss_and_grads("xdeepfm", p64, batch)
    m, params = T._build("xdeepfm", spec, cuda, cin_precision=prec)
    m.load_state(p64)
    from recsys_b200.xdeepfm import xdeepfm
    sp = xdeepfm.model_fn(T._features_to_torch(feats), batch["labels"], "train", params)
    m.backward(m.last["loss"])
    dg = m.dense_grads()
    print("==", prec, "loss", float(sp.loss), float(out64["loss"]))
    for k, g in g64.items():
        if k in dg:
            a = dg[k].detach().cpu().double().reshape(-1); b = g.reshape(-1)
            print("  %-16s err %.3e scale %.3e  gpu[:3]=%s ref[:3]=%s" % (k, float((a-b).abs().max()), float(b.abs().max()), a[:3].tolist(), b[:3].tolist()))
