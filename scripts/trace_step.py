"""Kernel timeline of graph-replayed train steps via torch.profiler (CUPTI): per-kernel device
time inside the replay and the idle gaps between consecutive kernels."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="deepfm")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--timeline", action="store_true",
                    help="also print the last step's records with start offsets and streams")
    a = ap.parse_args()
    args = bench.parse(["--model", a.model] + (["--batch", str(a.batch)] if a.batch else []))
    if args.batch is None:
        args.batch = 8192 if args.model == "xdeepfm" else 4096
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=int(os.environ.get("CTR_MAIN_PRIO", "-1"))))
    from recsys_b200 import ops
    from recsys_b200 import feature_column as fc
    from recsys_b200.data import SyntheticCriteo
    from recsys_b200.estimator import GraphedTrainStep
    if args.model == "din":
        args.n_batches = 8
        mod, params, host = bench.build_din(args, dev)
        sp = mod.model_fn({k: v.to(dev) for k, v in host[0][0].items()}, host[0][1].to(dev), "train", params)
    else:
        mod, params = bench.build_model(args, dev)
        lay = fc.layout(params["embedding_feature_columns"])
        host = SyntheticCriteo(lay, args.batch, 8, dist="uniform", seed=0, device=None).batches
        sp = mod.model_fn(ops.PackedFeatures(host[0][0].cont.to(dev), host[0][0].cat.to(dev),
                                             host[0][0].cont_keys, host[0][0].cat_keys),
                          host[0][1].to(dev), "train", params)
    sp.train_op()
    step = GraphedTrainStep(mod.model_fn, params, host[0][0], host[0][1], warmup=3)
    for i in range(5):
        step(*host[i % 8])
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(a.steps):
            step(*host[i % 8])
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"]
          if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    if not ev:
        print("no device activity records (CUPTI unavailable?)")
        return
    # split into steps at the blob D2D copy (first memcpy DtoD of each step)
    per = {}
    order = []
    last_end = None
    gaps = []
    t_first, t_last = ev[0]["ts"], ev[-1]["ts"] + ev[-1]["dur"]
    for e in ev:
        name = e["name"][:70]
        if name not in per:
            per[name] = [0.0, 0]
            order.append(name)
        per[name][0] += e["dur"]
        per[name][1] += 1
        if last_end is not None:
            gaps.append(max(0.0, e["ts"] - last_end))
        last_end = max(last_end or 0, e["ts"] + e["dur"])
    n = a.steps
    print("%d steps, span %.1f us/step, busy %.1f us/step, idle gaps %.1f us/step (%d records)" % (
        n, (t_last - t_first) / n, sum(v[0] for v in per.values()) / n, sum(gaps) / n, len(ev)))
    for name in order:
        tot, cnt = per[name]
        print("  %7.2f us x %4.1f/step  %s" % (tot / cnt, cnt / n, name))
    if a.timeline:
        # the last step = records from the last blob copy (Memcpy DtoD) on
        first = ev[0]["name"] if "Memcpy" not in ev[0]["name"] else next(
            (e["name"] for e in ev if "Memcpy" not in e["name"] and "split_lo" not in e["name"]), ev[0]["name"])
        starts = [i for i, e in enumerate(ev) if e["name"] == first]
        i0 = max(0, (starts[-1] if starts else len(ev) - len(ev) // n) - 4)
        t0 = ev[i0]["ts"]
        print("timeline of the last step (start us, duration us, stream, name):")
        for e in ev[i0:]:
            print("  %8.2f %7.2f  s%-4s %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"),
                                            e["name"][:60]))


if __name__ == "__main__":
    main()
