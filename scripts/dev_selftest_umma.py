"""Development check of the tcgen05 plumbing (nws_selftest_umma) against float64.
usage: dev_selftest_umma.py [K swap]   (no args: every config in its own subprocess)"""
import subprocess
import sys

sys.path.insert(0, ".")


def one(K, swap):
    import torch
    from neural_waveshaping_synthesis_b200 import _lib
    lib = _lib.load_library()
    g = torch.Generator().manual_seed(K)
    A = (torch.rand(128, K, generator=g) * 2 - 1).cuda()
    B = (torch.rand(64, K, generator=g) * 0.2 - 0.1).cuda()
    D = torch.zeros(128, 64, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.nws_selftest_umma(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, swap, st.data_ptr(), None)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    print("K=%3d swap=%d rc=%d status=%d max|err| vs fp64 = %.3e  (|ref|max %.3f)" % (K, swap, rc, int(st[0]), err, ref.abs().max().item()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for K in (8, 16, 104):
            for swap in (0, 1):
                r = subprocess.run([sys.executable, __file__, str(K), str(swap)], capture_output=True, text=True, timeout=120)
                out = (r.stdout + r.stderr).strip().splitlines()
                print("\n".join(l for l in out if "K=" in l or "rror" in l)[:600] or "(no output) rc=%d" % r.returncode, flush=True)
