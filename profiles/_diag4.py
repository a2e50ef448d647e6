import sys, time, torch
sys.path.insert(0, ".")
from envidr_b200 import render, scene
dev = torch.device("cuda:0")
t0 = time.time()
fp, bits, info = scene.fit_synthetic_field(0, device=dev, steps=1000, hidden_dim_env=256, ide_degree=5)
torch.cuda.synchronize(); print("fit", time.time() - t0, info)
cfg = render.RenderConfig(indir_ref=True)
for th in (40, 130):
    ro, rd = scene.camera_rays(800, 800, theta_deg=float(th)); ro, rd = ro.to(dev), rd.to(dev)
    for _ in range(3):
        st = []
        out = render.render(fp, bits, ro, rd, cfg, bg_color=1.0, stats=st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); render.render(fp, bits, ro, rd, cfg, bg_color=1.0); b.record(); torch.cuda.synchronize()
    print(th, "ms", a.elapsed_time(b), [(s.get("iterations"), s.get("samples")) for s in st], "ws mean", float(out["weights_sum"].mean()), "finite", bool(torch.isfinite(out["image"]).all()))
