"""Developer probe: a few extreme shapes (1 ... 4096 antennas, 1 ... 2000 beams, 16 ... 4096 samples) through dcbf_fused and
dcbf_beamform against float64 references.    python tests/probes/extreme_shapes.py
"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dpdk_dc_sand_b200 import _capi
from oracle import beamform_oracle as orc
TS = orc.SAMPLE_PERIOD
dev = torch.device("cuda", 0)
for (B, A, C, T, M) in [(1, 1000, 2, 64, 300), (1, 4096, 1, 32, 4), (1, 8, 2, 64, 2000), (1, 2, 3, 16, 1), (2, 1, 1, 16, 1), (1, 513, 2, 256, 3), (1, 64, 1, 4096, 16), (1, 600, 1, 256, 600)]:
    x = orc.make_samples(B, A, C, T, seed=A + M)
    dv = orc.make_delay_vals_random(C, M, A, seed=A * 3 + M)
    out = torch.full((B, 2, C, T // 16, 16, 2 * M), float("nan"), dtype=torch.float32, device=dev)
    try:
        _capi.fused(torch.from_numpy(x).to(dev), torch.from_numpy(dv).to(dev), out, B, A, C, C, T, M, 0, TS)
        torch.cuda.synchronize(); _capi.fused_status()
    except Exception as e:
        print((B, A, C, T, M), "EXC", e); continue
    ref = orc.beamform_pipeline(x, dv, C, 0, TS)
    budget = 2.0 ** -10 * orc.beamform_abs_bound(orc.reorder(x))[..., None]
    got = out.cpu().numpy()
    print((B, A, C, T, M), "nan" if np.isnan(got).any() else f"err/budget {np.max(np.abs(got - ref) / budget):.2e}", flush=True)
    re = torch.from_numpy(orc.reorder(x)).to(dev)
    if True:
        co = torch.randn((B, 2, C, 2 * A, 2 * M), dtype=torch.float32, device=dev)
        ob = torch.full_like(out, float("nan"))
        _capi.beamform(re, co, ob, B, C, T, A, M)
        torch.cuda.synchronize(); _capi.fused_status()
        xr = re.double().reshape(B, 2, C, T, 2 * A)
        refb = torch.matmul(xr, co.double()); scale = torch.matmul(xr, co.double().abs()) + 1e-30
        print("   beamform", "nan" if torch.isnan(ob).any() else f"{float(((ob.reshape(B,2,C,T,2*M).double()-refb).abs()/scale).max()):.2e}", flush=True)
