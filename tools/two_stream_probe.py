"""How much throughput do two compute streams (two depth maps in flight) add over one?  python tools/two_stream_probe.py [--workload dtu]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="dtu"); ap.add_argument("--iters", type=int, default=12)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda:0")
net, _ = bench.make_net(); net = net.to(dev)
ins = []
for it in range(2):
    feats, proj, dv = bench.make_inputs(wl, 1234 + it)
    ins.append(({k: v.to(dev) for k, v in feats.items()}, {k: v.to(dev) for k, v in proj.items()}, dv.to(dev)))
def run(nstreams, iters):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    outs = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    for i in range(iters):
        with torch.cuda.stream(streams[i % nstreams]):
            f, p, d = ins[i % 2]
            outs.append(net.forward_features(f, p, d, bench.TMP)["refined_depth"])
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, outs
for n in (1, 2, 1, 2, 3):
    run(n, 4)
    t0 = time.time(); ms, outs = run(n, a.iters); wall = (time.time() - t0) * 1e3 / a.iters
    print(f"streams {n}: {ms:.3f} ms per depth map ({1000/ms:.1f} maps/s), host wall {wall:.3f} ms")
ms1, o1 = run(1, 2); ms2, o2 = run(2, 2)
print("same results:", all(torch.equal(a_, b_) for a_, b_ in zip(o1, o2)))
