"""Diagnostic (not part of the product): CPU-side timeline of bench-like steps to find host stalls."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gsgen_b200.scenes import make_scene
from gsgen_b200.parallel import ViewParallelRenderer
from gsgen_b200.rasterizer import render_view

print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
print("torch threads", torch.get_num_threads(), "OMP", os.environ.get("OMP_NUM_THREADS"))
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if nthreads: torch.set_num_threads(nthreads)
dev = torch.device("cuda:0")
sc = make_scene(wl)
C = sc.C
cam, c2w = sc.cams[0], sc.c2ws[0]
vpr = ViewParallelRenderer(dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, alpha=sc.alpha, sh=sc.sh), C, dev)
gout = torch.randn(cam.h, cam.w, 3, device=dev)
def step(trace):
    t0 = time.perf_counter(); vpr.zero_grad(); t1 = time.perf_counter()
    out = render_view(vpr.params["mean"], vpr.params["qvec"], vpr.params["svec"], vpr.params["alpha"], c2w, cam, sh=vpr.params["sh"], C=C, grad_sink=vpr.grad_views)
    t2 = time.perf_counter(); out["rgb"].backward(gradient=gout); t3 = time.perf_counter()
    trace.append((t1 - t0, t2 - t1, t3 - t2))
for rep in range(4):
    tr = []
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(100): step(tr)
    e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
    z = [sum(x) for x in tr]
    print(f"rep {rep}: gpu-event {e0.elapsed_time(e1)/100:.3f} ms/step wall {(w1-w0)*10:.3f} ms/step | cpu per step: zero {statistics.mean(x[0] for x in tr)*1e3:.3f} fwd {statistics.mean(x[1] for x in tr)*1e3:.3f} bwd {statistics.mean(x[2] for x in tr)*1e3:.3f} | max step {max(z)*1e3:.2f} ms, p50 {statistics.median(z)*1e3:.3f}")
    slow = sorted(range(100), key=lambda i: -z[i])[:5]
    print("   slowest steps:", [(i, round(tr[i][0]*1e3,2), round(tr[i][1]*1e3,2), round(tr[i][2]*1e3,2)) for i in slow])
