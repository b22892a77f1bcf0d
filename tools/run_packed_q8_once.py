import sys
import torch
sys.path.insert(0, "/root/repo")
from dpdk_dc_sand_b200 import _capi
TS = 1 / 1712e6
A, C, T, M, B = 64, 4096, 256, 64, 1
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
gains = torch.full((M,), 0.01, dtype=torch.float32, device=dev)
out = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.int8, device=dev)
pk = torch.empty(_capi.fused_packed_bytes(A, C, M), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
_capi.fused_pack_coeffs_q8(dv, gains, pk, A, C, C, M, 0, TS)
for _ in range(4):
    _capi.fused_packed_q8(x, pk, gains, out, B, A, C, C, T, M, 0, TS)
torch.cuda.synchronize()
_capi.fused_status()
print("ok")
