#!/usr/bin/env python
"""Where the step's wall time goes BETWEEN kernels: runs bench.py's device-resident step under torch.profiler (CUPTI kernel
records: start / duration per launch - nsys is not in this image), then prints kernel time, idle time on the stream, and the
kernel pairs with the largest gaps.   python tools/step_gaps.py [--workload visual_bert] [--steps 2] > profiles/..."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="visual_bert")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    wl = bench.WORKLOADS[args.workload]()
    B = args.batch or wl.default_batch
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = wl.build(0.1).to(dev).train()
    batch = bench.to_device(wl.host_batch(B, 1), dev)
    aux = wl.aux(B, dev)

    def step():
        net.zero_grad(set_to_none=True)
        loss = wl.loss(net, batch, aux)
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # host synchronisations inside the step (each one drains the launch queue): torch reports them with a stack
    import warnings
    torch.cuda.set_sync_debug_mode("warn")
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        step()
    torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    print("synchronising calls in one step: %d" % len(caught))
    seen = set()
    for w in caught:
        key = (w.filename, w.lineno)
        if key not in seen:
            seen.add(key)
            print("  %s:%d  %s" % (w.filename.replace(ROOT + "/", ""), w.lineno, str(w.message)[:100]))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "mmfb_step_trace.json")
    prof.export_chrome_trace(path)
    with open(path) as fh:
        ev = [e for e in json.load(fh)["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    busy = sum(e["dur"] for e in ev)
    gaps = []
    end = ev[0]["ts"] + ev[0]["dur"]
    for a, b in zip(ev, ev[1:]):
        g = b["ts"] - end
        if g > 0:
            gaps.append((g, a["name"][:60], b["name"][:60]))
        end = max(end, b["ts"] + b["dur"])
    idle = sum(g for g, _, _ in gaps)
    n = args.steps
    print("%s B=%d: %d device activities over %d steps; per step: wall %.3f ms, kernels %.3f ms, idle %.3f ms (%.1f %%), %d launches"
          % (args.workload, B, len(ev), n, (t1 - t0) / n / 1e3, busy / n / 1e3, idle / n / 1e3, 100.0 * idle / (t1 - t0), len(ev) // n))
    hist = {}
    for g, a, b in gaps:
        k = (a.split("(")[0].split("<")[0][-40:], b.split("(")[0].split("<")[0][-40:])
        hist.setdefault(k, [0, 0.0])
        hist[k][0] += 1
        hist[k][1] += g
    print("gap histogram: <2us %d, 2-5us %d, 5-20us %d, >20us %d" % (
        sum(1 for g, _, _ in gaps if g < 2), sum(1 for g, _, _ in gaps if 2 <= g < 5), sum(1 for g, _, _ in gaps if 5 <= g < 20),
        sum(1 for g, _, _ in gaps if g >= 20)))
    print("largest idle by (previous kernel -> next kernel), us per step:")
    for (a, b), (c, t) in sorted(hist.items(), key=lambda kv: -kv[1][1])[:25]:
        print("  %8.1f us  n=%3d  avg %6.1f   %s -> %s" % (t / n, c // n, t / c, a, b))


if __name__ == "__main__":
    main()
