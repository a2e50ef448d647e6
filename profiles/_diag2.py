import sys, torch
sys.path.insert(0, ".")
from envidr_b200 import dist as D, render, scene
ws = int(sys.argv[1]); W = 1600
dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5); fp.precision = "tc"; fp = fp.to(dev).pack()
bf = torch.from_numpy(scene.make_bitfield()).to(dev)
ro, rd = scene.camera_rays(W, W); ro, rd = ro.to(dev), rd.to(dev)
cfg = render.RenderConfig(indir_ref=True)
idx = D.tile_shard_indices(W, W, 0, ws).to(dev)
o, d = ro[idx].contiguous(), rd[idx].contiguous()
for _ in range(3):
    render.render(fp, bf, o, d, cfg, bg_color=1.0)
torch.cuda.synchronize()
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
flush.zero_(); torch.cuda.synchronize()
render.render(fp, bf, o, d, cfg, bg_color=1.0)
torch.cuda.synchronize()
flush.zero_(); torch.cuda.synchronize()
