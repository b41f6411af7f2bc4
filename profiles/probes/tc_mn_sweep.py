"""Sweep MN-major UMMA descriptor conventions (LBO, SBO, K-step bytes) for the tcgen05 GEMM."""
import itertools, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import numpy as np, torch, importlib
    dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
    M, N, K = 256, 384, 64
    rng = np.random.default_rng(0)
    A = (rng.standard_normal((M, K)) / 8).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    for ta, tb in ((0, 1), (1, 0), (1, 1)):
        a = torch.from_numpy(np.ascontiguousarray(A.T if ta else A)).cuda(); b = torch.from_numpy(np.ascontiguousarray(B.T if tb else B)).cuda()
        c = torch.zeros(M, N, device="cuda")
        rc = lib.dvae_tc_linear(L.ptr(a), a.stride(0), ta, L.ptr(b), b.stride(0), tb, L.ptr(c), N, M, N, K, None, None, 0.0, 0, 3, L.stream_ptr())
        try:
            torch.cuda.synchronize()
            err = np.abs(c.cpu().numpy() - want).max() / np.abs(want).max()
            frac0 = float((c == 0).float().mean())
        except Exception as e:
            err, frac0 = str(e)[:60], -1
        print(os.environ.get("DVAE_TC_MN"), (ta, tb), "rc", rc, "relerr", err, "zeros", frac0, flush=True)
else:
    for lbo, sbo, ks in ((4096, 512, 1024), (4096, 1024, 1024), (4096, 512, 512), (16, 512, 1024), (512, 4096, 1024)):
        env = dict(os.environ, DVAE_TC_MN=f"{lbo},{sbo},{ks}")
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=120)
        print(r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
