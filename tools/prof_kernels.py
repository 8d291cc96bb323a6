"""One launch of each heavy kernel at 2^18 points (for ncu): python tools/prof_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kzg_setup_powersoftau_b200 as kz
ctx = kz.Context(1)
N = 1 << int(os.environ.get("LOGN", "18"))
tau = 0x1234567890ABCDEF1234567890ABCDEF
status = torch.full((1,), -1, dtype=torch.int64, device="cuda")
L = kz._ffi.lib()
KINDS = ((1, 1, 3, 14), (1, 2, 3, 14), (2, 1, 3, 14), (2, 2, 3, 14), (2, 2, 1, 0), (1, 3, 4, 0))
if os.environ.get("ONLY_G2"):  # the two G2 kernels with the subgroup ladder only (a short capture)
    KINDS = ((2, 1, 3, 14), (2, 2, 3, 14))
for group, in_fmt, out_fmt, checks in KINDS:
    ri, ro = L.ptau_record_size(group, in_fmt), L.ptau_record_size(group, out_fmt)
    d_in = torch.empty(N * ri, dtype=torch.uint8, device="cuda"); d_out = torch.empty(N * ro, dtype=torch.uint8, device="cuda")
    gen_fmt = 1 if in_fmt == 3 else in_fmt
    ctx.generate_device(group, gen_fmt, 1, tau, 0, N, d_in.data_ptr()); torch.cuda.synchronize()
    if in_fmt == 3:
        t = torch.empty_like(d_in); ctx.convert_device(group, 1, d_in.data_ptr(), 3, t.data_ptr(), N, 0, status.data_ptr()); torch.cuda.synchronize(); d_in = t
    for _ in range(2):
        ctx.convert_device(group, in_fmt, d_in.data_ptr(), out_fmt, d_out.data_ptr(), N, checks, status.data_ptr())
    torch.cuda.synchronize()
    assert int(status.item()) == -1
print("ok")
