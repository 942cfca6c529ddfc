import ctypes as C, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sphy_b200 import _lib
from sphy_b200.model import CatchmentSource, SphyModel
from oracle.ldd_oracle import LddStruct
size = 10000
cat = bench.make_case(size)
tmp = tempfile.mkdtemp()
cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
a = SphyModel(cfg, CatchmentSource(cat), route_method="scan", collect_tss=False)
a.initial()
n = a.n
lib = a.lib
g = torch.Generator(device="cuda").manual_seed(3)
mm = (8.0 * torch.rand(n, device="cuda", generator=g) * (torch.rand(n, device="cuda", generator=g) < 0.6)).float().contiguous()
q_lv = torch.zeros(n, device="cuda"); q_sc = torch.zeros(n, device="cuda")
src = (C.c_void_p * 1)(mm.data_ptr())
kx = _lib.SphyParam(None, 0.0)
for q, method in ((q_lv, _lib.ROUTE_LEVELS), (q_sc, _lib.ROUTE_SCAN)):
    dst = (C.c_void_p * 1)(q.data_ptr())
    _lib.check(lib, a._ctx, lib.sphy_route_step(a._ctx, 1, src, dst, kx, C.c_float(1.0), method, None), "route")
torch.cuda.synchronize()
rel = ((q_sc.double() - q_lv.double()).abs() / q_lv.double().abs().clamp_min(1e-12))
print("scan vs levels: max", rel.max().item(), "cells > 1e-5:", int((rel > 1e-5).sum().item()), "> 5e-6:", int((rel > 5e-6).sum().item()), "of", n)
t0 = time.time()
st = LddStruct(np.asarray(cat.ldd, np.uint8))
print("cpu ldd build", time.time() - t0)
canon = a.cell_order
mm_c = np.empty(n, np.float32); mm_c[canon] = mm.cpu().numpy()
rr = (mm_c * np.float32(0.001) * np.float32(250.0 * 250.0)) / np.float32(86400)
t0 = time.time(); f64 = st.accuflux_f64(rr); print("cpu f64 accuflux", time.time() - t0)
f32o = st.accuflux(rr)
lv = np.empty(n, np.float64); lv[canon] = q_lv.cpu().numpy()
sc = np.empty(n, np.float64); sc[canon] = q_sc.cpu().numpy()
den = np.maximum(np.abs(f64), 1e-12)
print("levels bit-equal to REAL4 oracle:", bool(np.array_equal(lv.astype(np.float32), f32o)))
print("levels vs float64 truth: max", float(np.max(np.abs(lv - f64) / den)))
print("scan   vs float64 truth: max", float(np.max(np.abs(sc - f64) / den)))
