import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_estimation_b200 import StereoUNet
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = StereoUNet().to(dev).eval()
x = torch.rand(1, 6, 240, 320, device=dev)
with torch.inference_mode():
    for _ in range(5): model(x, return_uncertainty=True)
    os.environ["SDN_CUDA_GRAPH"] = "0"
    model.profile_enable(True)
    for _ in range(3): model(x, return_uncertainty=True)
    rows = model.profile_dump(); model.profile_enable(False)
    os.environ["SDN_CUDA_GRAPH"] = "1"
    names=[f"{bk}.{i}" for bk in ["enc1","enc2","enc3","enc4","bott","dec4","dec3","dec2","dec1"] for i in (0,3)]
    print("per-op us (eager, B=1):", ' '.join(f"{r['name'][:6]}{names[r['layer']] if r['layer']<18 else r['layer']}:{r['ms']/3*1e3:.0f}" for r in rows))
    print("sum us", sum(r['ms'] for r in rows)/3*1e3)
    for _ in range(10): model(x, return_uncertainty=True)
    ts=[]
    for _ in range(200):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); model(x, return_uncertainty=True); b.record(); b.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    print("graph call p50 us", np.percentile(ts,50), "p99", np.percentile(ts,99))
    graph = model._engine.graphs[(1, True)][0]
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): graph.replay()
    b.record(); b.synchronize()
    print("graph replay back-to-back us each", a.elapsed_time(b)*1e3/50)
    import time
    t0=time.perf_counter()
    for _ in range(200): model(x, return_uncertainty=True)
    torch.cuda.synchronize(); print("wall us per call (pipelined)", (time.perf_counter()-t0)/200*1e6)
