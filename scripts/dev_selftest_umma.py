"""Development check of the tcgen05 plumbing (nws_selftest_umma) against float64."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from neural_waveshaping_synthesis_b200 import _lib  # noqa: E402

lib = _lib.load_library()
for K in (8, 16, 104):
    for swap in (0, 1):
        g = torch.Generator().manual_seed(K)
        A = (torch.rand(128, K, generator=g) * 2 - 1).cuda()
        B = (torch.rand(64, K, generator=g) * 0.2 - 0.1).cuda()
        D = torch.zeros(128, 64, device="cuda")
        st = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = lib.nws_selftest_umma(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, swap, st.data_ptr(), None)
        torch.cuda.synchronize()
        ref = A.double() @ B.double().t()
        err = (D.double() - ref).abs().max().item()
        tf32 = (A.double() @ B.double().t() - (A @ B.t()).double()).abs().max().item()
        print("K=%3d swap=%d rc=%d status=%d max|err| vs fp64 = %.3e (torch fp32 matmul: %.3e)" % (K, swap, rc, int(st[0]), err, tf32))
