"""CPU simulation (torch.float8_e4m3fn vs float64) of the mode-1 arithmetic of csrc/field_tc.cu: fp16 main product + e4m3 correction products.
Run: python profiles/fp8_correction_sim.py   (no GPU needed; results quoted in the kernel header and DESIGN.md)"""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from envidr_b200 import scene
from oracle import oracle as O
torch.manual_seed(0)
fp = scene.make_synthetic_field(0)
M = 20000
g = torch.Generator().manual_seed(1)
d = torch.nn.functional.normalize(torch.randn(M,3,generator=g,dtype=torch.float64),dim=-1)
rough = torch.rand(M,1,generator=g,dtype=torch.float64)*0.3
x0 = O.ide_encode(d, rough, 5).double()          # [M,72]
print("ide", x0.shape, x0.abs().max())
env = [(W.double(), b.double()) for W,b in fp.env]
color = [(W.double(), b.double()) for W,b in fp.color]
def f16(x): return x.float().half().double()
def split16(x):
    hi = f16(x); lo = f16(x - hi); return hi, lo
def e4m3(x):
    return x.float().clamp(-448,448).to(torch.float8_e4m3fn).double()
def e5m2(x):
    return x.float().clamp(-57344,57344).to(torch.float8_e5m2).double()
def layer(x, W, b, scheme):
    if scheme == "exact":
        return x @ W.t() + b
    xh, xl = split16(x); wh, wl = split16(W)
    if scheme == "s3":
        return (xh @ wh.t() + xl @ wh.t() + xh @ wl.t()).float().double() + b
    if scheme == "s1":
        return (xh @ wh.t()).float().double() + b
    if scheme == "s8":       # corrections in e4m3
        SA, SW = 2.0**11, 2.0**4
        c = e4m3(xl*SA) @ e4m3(wh*SW).t() + e4m3(xh) @ e4m3(wl*2.0**15).t()
        return (xh @ wh.t() + c * 2.0**-15).float().double() + b
    if scheme == "s8b":      # hi operand of corrections in e4m3, lo in e4m3; weights hi scaled per-tensor
        SA, SW = 2.0**11, 2.0**4
        c = e4m3(xl*SA) @ e4m3(wh*SW).t() + e4m3(xh) @ e4m3(wl*2.0**15).t()
        return (xh @ wh.t() + c * 2.0**-15).float().double() + b
    raise ValueError
def envnet(x, scheme, first="s3"):
    for i,(W,b) in enumerate(env):
        sch = scheme if (i>0 or scheme in ("exact",)) else first
        if scheme=="exact": sch="exact"
        x = layer(x, W, b, sch)
        if i < len(env)-1: x = torch.relu(x)
    return x / x.norm(dim=-1,keepdim=True).clamp_min(1e-12)
ref = envnet(x0, "exact")
geo = torch.nn.functional.normalize(torch.randn(M,12,generator=g,dtype=torch.float64),dim=-1)
nrm = torch.nn.functional.normalize(torch.randn(M,3,generator=g,dtype=torch.float64),dim=-1)
ndw = torch.rand(M,1,generator=g,dtype=torch.float64)
def colornet(f):
    x = torch.cat([geo,nrm,f,ndw],-1)
    for i,(W,b) in enumerate(color):
        x = x @ W.t() + b
        if i < len(color)-1: x = torch.relu(x)
    return torch.sigmoid(x)
rgb_ref = colornet(ref)
for scheme, first in (("s3","s3"),("s8","s3"),("s8","s8"),("s1","s1")):
    f = envnet(x0, scheme, first)
    e = (f-ref).abs().max(-1).values
    er = (colornet(f)-rgb_ref).abs().max(-1).values
    print(f"{scheme}/{first}: feat max {e.max():.3e} p99.9 {e.quantile(0.999):.3e} mean {e.mean():.3e} | rgb max {er.max():.3e} p99.9 {er.quantile(0.999):.3e} mean {er.mean():.3e}")
# activation ranges
x = x0
for i,(W,b) in enumerate(env[:-1]):
    x = torch.relu(x @ W.t() + b); print("layer",i,"act max",float(x.max()),"mean",float(x.mean()))

print("---- scaled scheme + trained weights (relight_mlps.npz: env 160 / deg 4, shipped color net)")
z = np.load(os.path.join(ROOT, "tests", "golden", "relight_mlps.npz"))
n = lambda name: sorted({int(k.split("_")[2]) for k in z.files if k.startswith(name + "_")})
L = lambda name: [(torch.from_numpy(z[f"{name}_{i}_weight"]).double(), torch.from_numpy(z[f"{name}_{i}_bias"]).double()) for i in n(name)]
env_t, color_t = L("env_net"), L("color_net")
print([tuple(W.shape) for W,_ in env_t], [float(W.abs().max()) for W,_ in env_t])
x0t = O.ide_encode(d, rough, 4).double()
def layer2(x, W, b, scheme, SAH=3, SWH=4):
    if scheme == "exact": return x @ W.t() + b
    xh, xl = split16(x); wh, wl = split16(W)
    if scheme == "s3": return (xh @ wh.t() + xl @ wh.t() + xh @ wl.t()).float().double() + b
    SM = SAH + SWH + 11 - 15          # scale of the main product so that scale-input-d = 15 aligns the corrections
    c = e4m3(xl*2.0**(SAH+11)) @ e4m3(wh*2.0**SWH).t() + e4m3(xh*2.0**SAH) @ e4m3(wl*2.0**(SWH+11)).t()
    main = xh @ f16(wh*2.0**SM).t()
    return ((main + c*2.0**-15).float().double()) * 2.0**-SM + b
def net(x, layers, scheme, first, **kw):
    for i,(W,b) in enumerate(layers):
        sch = "exact" if scheme=="exact" else (first if i==0 else scheme)
        x = layer2(x, W, b, sch, **kw)
        if i < len(layers)-1: x = torch.relu(x)
    return x
def unit(x): return x / x.norm(dim=-1,keepdim=True).clamp_min(1e-12)
for name, envL, colL, xin in (("xavier256", env, color, x0), ("trained160", env_t, color_t, x0t)):
    ref = unit(net(xin, envL, "exact", None))
    def col(f):
        x = torch.cat([geo,nrm,f,ndw],-1)
        return torch.sigmoid(net(x, colL, "exact", None))
    rr = col(ref)
    acts = xin
    for i,(W,b) in enumerate(envL[:-1]):
        acts = torch.relu(acts @ W.t() + b); print(name, "layer", i, "act max", float(acts.max()), "mean", float(acts.mean()))
    for scheme, first, kw in (("s3","s3",{}),("s8","s3",dict(SAH=3,SWH=4)),("s8","s8",dict(SAH=3,SWH=4)),("s8","s3",dict(SAH=0,SWH=4)),("s8","s3",dict(SAH=5,SWH=2))):
        f = unit(net(xin, envL, scheme, first, **kw))
        e = (f-ref).abs().max(-1).values; er = (col(f)-rr).abs().max(-1).values
        print(f"{name} {scheme}/{first} {kw}: feat max {e.max():.3e} mean {e.mean():.3e} | rgb max {er.max():.3e} p99.9 {er.quantile(0.999):.3e} mean {er.mean():.3e}")
