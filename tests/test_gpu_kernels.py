"""Kernel-level parity: every C-ABI entry point against the numpy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): fp32 forward within 1e-5 relative, gradients within 1e-3
relative; integer outputs (argmax) identical.  The oracle is evaluated in float64.
"""
import os

import numpy as np
import pytest
import torch

from oracle import dvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(dvae):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return dvae._lib.load()


@pytest.fixture(scope="module")
def L(dvae):
    return dvae._lib


_KEEP = []     # device tensors stay alive until the test module is done: the C ABI takes raw pointers


def dev(x, dtype=torch.float32):
    t = torch.as_tensor(np.ascontiguousarray(x)).to(device="cuda", dtype=dtype).contiguous()
    _KEEP.append(t)
    if len(_KEEP) > 4096:
        torch.cuda.synchronize()
        del _KEEP[:2048]
    return t


def rel(a, b):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(5, 7, 3), (64, 64, 16), (130, 257, 100), (300, 1024, 256), (2560, 1024, 256), (17, 40, 1030)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_linear(lib, L, M, N, K, ta, tb):
    rng = np.random.default_rng(M * 7 + N * 3 + K + ta * 2 + tb)
    A = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    Bm = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    a_d = dev(A.T if ta else A)
    b_d = dev(Bm.T if tb else Bm)
    c_d = dev(C0)
    st = L.stream_ptr()
    L.check(lib.dvae_linear(L.ptr(a_d), a_d.stride(0), ta, L.ptr(b_d), b_d.stride(0), tb, L.ptr(c_d), N, M, N, K,
                            L.ptr(dev(bias)), None, 1.0, 0, st), "linear")
    want = A.astype(np.float64) @ Bm.astype(np.float64).T + bias + C0
    assert rel(c_d, want) < 2e-6
    L.check(lib.dvae_linear(L.ptr(a_d), a_d.stride(0), ta, L.ptr(b_d), b_d.stride(0), tb, L.ptr(c_d), N, M, N, K,
                            L.ptr(dev(bias)), L.ptr(dev(bias)), 0.0, 1, st), "linear")
    want = np.tanh(A.astype(np.float64) @ Bm.astype(np.float64).T + 2 * bias)
    assert np.abs(c_d.cpu().numpy() - want).max() < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 64), (130, 257, 100), (2688, 1024, 256), (100, 10000, 256), (1024, 256, 2688)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_tc_linear_3xtf32(lib, L, M, N, K, ta, tb):
    """tcgen05 / TMA / TMEM GEMM: 3xTF32 must be fp32-grade, 1xTF32 within TF32 rounding."""
    rng = np.random.default_rng(M + 3 * N + 7 * K + ta * 2 + tb)
    A = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    Bm = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    # TMA needs row strides that are multiples of 16 bytes: pad the leading dimension of the stored matrix
    def stored(X, t):
        S = np.ascontiguousarray(X.T if t else X)
        ld = (S.shape[1] + 3) // 4 * 4
        buf = torch.zeros(S.shape[0], ld, device="cuda")
        buf[:, :S.shape[1]] = torch.from_numpy(S).cuda()
        _KEEP.append(buf)
        return buf, ld
    a_d, lda = stored(A, ta)
    b_d, ldb = stored(Bm, tb)
    want = A.astype(np.float64) @ Bm.astype(np.float64).T + bias
    bias_d = dev(bias)
    st = L.stream_ptr()
    for passes, tol in ((3, max(2e-6, 1e-8 * K)), (1, 2e-3)):   # in-TMEM fp32 accumulation error grows ~5e-9*K
        c_d = torch.full((M, N), 7.0, device="cuda")
        L.check(lib.dvae_tc_linear(L.ptr(a_d), lda, ta, L.ptr(b_d), ldb, tb, L.ptr(c_d), N, M, N, K, L.ptr(bias_d), None,
                                   0.0, 0, passes, st), "tc_linear")
        torch.cuda.synchronize()
        assert rel(c_d, want) < tol, (passes, rel(c_d, want))
    # beta accumulate + second bias
    c_d = torch.ones(M, N, device="cuda")
    L.check(lib.dvae_tc_linear(L.ptr(a_d), lda, ta, L.ptr(b_d), ldb, tb, L.ptr(c_d), N, M, N, K, L.ptr(bias_d), L.ptr(bias_d),
                               1.0, 0, 3, st), "tc_linear")
    assert rel(c_d, want + bias + 1.0) < max(2e-6, 1e-8 * K)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 64), (130, 257, 100), (2688, 1024, 256), (100, 10000, 256),
                                   (1024, 256, 2688), (2816, 256, 1024), (64, 64, 36)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_tc16_linear_fp16_split(lib, L, M, N, K, ta, tb):
    """fp16 (hi, lo)-split tcgen05 GEMM: fp32-grade against float64, incl. transposed operands, partial tiles, split-K,
    multi-tile CTAs and the operand scales (host constant and device amax) on tiny-magnitude 'gradient' operands."""
    rng = np.random.default_rng(M + 3 * N + 7 * K + ta * 2 + tb)
    A = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    Bm = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)

    def stored(X, t):
        S = np.ascontiguousarray(X.T if t else X)
        ld = (S.shape[1] + 3) // 4 * 4
        buf = torch.zeros(S.shape[0], ld, device="cuda")
        buf[:, :S.shape[1]] = torch.from_numpy(S).cuda()
        _KEEP.append(buf)
        return buf, ld
    a_d, lda = stored(A, ta)
    b_d, ldb = stored(Bm, tb)
    want = A.astype(np.float64) @ Bm.astype(np.float64).T + bias
    bias_d = dev(bias)
    st = L.stream_ptr()
    tol = max(2e-6, 1e-8 * K)
    c_d = torch.full((M, N), 7.0, device="cuda")
    L.check(lib.dvae_tc16_linear(L.ptr(a_d), lda, ta, L.ptr(b_d), ldb, tb, L.ptr(c_d), N, M, N, K, L.ptr(bias_d), None,
                                 0.0, 0, 1.0, 1.0, None, None, st), "tc16_linear")
    assert rel(c_d, want) < tol, rel(c_d, want)
    # beta accumulate + second bias
    c_d = torch.ones(M, N, device="cuda")
    L.check(lib.dvae_tc16_linear(L.ptr(a_d), lda, ta, L.ptr(b_d), ldb, tb, L.ptr(c_d), N, M, N, K, L.ptr(bias_d), L.ptr(bias_d),
                                 1.0, 0, 1.0, 1.0, None, None, st), "tc16_linear")
    assert rel(c_d, want + bias + 1.0) < tol
    # a "gradient" A operand far below fp16's normal range: exact with the measured-amax scale, and with a host scale
    tiny = 1e-9
    a_t, _ = stored(A * np.float32(tiny), ta)
    amax = torch.tensor([float(np.abs(A * np.float32(tiny)).max())], device="cuda").view(torch.int32)
    want_t = (A * np.float32(tiny)).astype(np.float64) @ Bm.astype(np.float64).T
    c_d = torch.zeros(M, N, device="cuda")
    L.check(lib.dvae_tc16_linear(L.ptr(a_t), lda, ta, L.ptr(b_d), ldb, tb, L.ptr(c_d), N, M, N, K, None, None,
                                 0.0, 0, 1.0, 1.0, L.ptr(amax), None, st), "tc16_linear")
    assert rel(c_d, want_t) < tol, rel(c_d, want_t)
    c_d = torch.zeros(M, N, device="cuda")
    L.check(lib.dvae_tc16_linear(L.ptr(a_t), lda, ta, L.ptr(b_d), ldb, tb, L.ptr(c_d), N, M, N, K, None, None,
                                 0.0, 0, float(2 ** 28), 1.0, None, None, st), "tc16_linear")
    assert rel(c_d, want_t) < tol, rel(c_d, want_t)


@pytest.mark.parametrize("M,R,C", [(300, 1024, 256), (2688, 1024, 512), (130, 384, 100)])
def test_tc16_linear_with_registered_weight_planes(lib, L, M, R, C):
    """Hybrid GEMM: B is a registered weight whose fp16 (hi, lo) planes were written once (dvae_weight_planes_refresh) and
    arrive by bulk copy; A still goes through the converter ring.  Both orientations (x . W^T and g . W), blocks of W that
    start on a 128-row / 32-column boundary, and the registry switched off again."""
    rng = np.random.default_rng(M + R)
    W = (rng.standard_normal((R, C)) / np.sqrt(C)).astype(np.float32)
    Wd = dev(W)
    pl = torch.empty(lib.dvae_weight_planes_floats(R, C, 0), device="cuda")
    plt = torch.empty(lib.dvae_weight_planes_floats(R, C, 1), device="cuda")
    st = L.stream_ptr()
    L.check(lib.dvae_weight_planes_clear(), "clear")
    L.check(lib.dvae_weight_planes_register(L.ptr(Wd), R, C, L.ptr(pl), L.ptr(plt)), "register")
    L.check(lib.dvae_weight_planes_refresh(st), "refresh")
    try:
        for enable in (1, 0):
            L.check(lib.dvae_weight_planes_enable(enable), "enable")
            # y = x . W^T + b   (trans_b = 0; N = R, K = C)
            x = rng.standard_normal((M, C)).astype(np.float32)
            b = rng.standard_normal(R).astype(np.float32)
            y = torch.zeros(M, R, device="cuda")
            L.check(lib.dvae_tc16_linear(L.ptr(dev(x)), C, 0, L.ptr(Wd), C, 0, L.ptr(y), R, M, R, C, L.ptr(dev(b)), None, 0.0, 0,
                                         1.0, 1.0, None, None, st), "x W^T")
            assert rel(y, x.astype(np.float64) @ W.astype(np.float64).T + b) < 3e-6
            # d = g . W         (trans_b = 1: B stored [K = R, N = C])
            g = rng.standard_normal((M, R)).astype(np.float32)
            d = torch.zeros(M, C, device="cuda")
            L.check(lib.dvae_tc16_linear(L.ptr(dev(g)), R, 0, L.ptr(Wd), C, 1, L.ptr(d), C, M, C, R, None, None, 0.0, 0,
                                         1.0, 1.0, None, None, st), "g W")
            assert rel(d, g.astype(np.float64) @ W.astype(np.float64)) < 3e-6
            if R >= 512:
                # a row block of W as the K range of g . W[r0:r0+K, :]  (the vocabulary chunks of d_h = P . W_out)
                r0, K = 256, R - 256 - 24
                g2 = rng.standard_normal((M, K)).astype(np.float32)
                d2 = torch.zeros(M, C, device="cuda")
                L.check(lib.dvae_tc16_linear(L.ptr(dev(g2)), K, 0, Wd.data_ptr() + 4 * r0 * C, C, 1, L.ptr(d2), C, M, C, K, None, None,
                                             0.0, 0, 1.0, 1.0, None, None, st), "g W[r0:]")
                assert rel(d2, g2.astype(np.float64) @ W[r0:r0 + K].astype(np.float64)) < 3e-6
                # a row block of W as the N range of x . W[r0:r0+N, :]^T
                y2 = torch.zeros(M, 256, device="cuda")
                L.check(lib.dvae_tc16_linear(L.ptr(dev(x)), C, 0, Wd.data_ptr() + 4 * 128 * C, C, 0, L.ptr(y2), 256, M, 256, C, None, None,
                                             0.0, 0, 1.0, 1.0, None, None, st), "x W[128:384]^T")
                assert rel(y2, x.astype(np.float64) @ W[128:384].astype(np.float64).T) < 3e-6
    finally:
        L.check(lib.dvae_weight_planes_clear(), "clear")


def test_tc16_linear_tanh_and_rejects_unaligned(lib, L):
    rng = np.random.default_rng(5)
    M, N, K = 128, 512, 64
    A = rng.standard_normal((M, K)).astype(np.float32) * 0.3
    Bm = rng.standard_normal((N, K)).astype(np.float32) * 0.3
    bias = rng.standard_normal(N).astype(np.float32)
    c_d = torch.zeros(M, N, device="cuda")
    L.check(lib.dvae_tc16_linear(L.ptr(dev(A)), K, 0, L.ptr(dev(Bm)), K, 0, L.ptr(c_d), N, M, N, K, L.ptr(dev(bias)), None,
                                 0.0, 1, 1.0, 1.0, None, None, L.stream_ptr()), "tc16_linear")
    assert rel(c_d, np.tanh(A.astype(np.float64) @ Bm.astype(np.float64).T + bias)) < 2e-6
    x = torch.zeros(64 * 70, device="cuda")
    rc = lib.dvae_tc16_linear(L.ptr(x), 70, 0, L.ptr(x), 70, 0, L.ptr(c_d), 64, 64, 64, 70, None, None, 0.0, 0, 1.0, 1.0, None, None, None)
    assert rc == -1 and b"aligned" in lib.dvae_last_error_string()


def test_linear_strided_output_and_colsum(lib, L):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((70, 33)).astype(np.float32)
    Bm = rng.standard_normal((20, 33)).astype(np.float32)
    big = torch.zeros(70, 50, device="cuda")
    L.check(lib.dvae_linear(L.ptr(dev(A)), 33, 0, L.ptr(dev(Bm)), 33, 0, big.data_ptr() + 4 * 10, 50, 70, 20, 33,
                            None, None, 0.0, 0, L.stream_ptr()), "linear")
    assert rel(big[:, 10:30], A.astype(np.float64) @ Bm.T) < 2e-6
    assert float(big[:, :10].abs().max()) == 0 and float(big[:, 30:].abs().max()) == 0
    out = torch.zeros(20, device="cuda")
    L.check(lib.dvae_colsum(big.data_ptr() + 40, 50, 70, 20, L.ptr(out), 0.0, L.stream_ptr()), "colsum")
    assert rel(out, (A.astype(np.float64) @ Bm.T).sum(0)) < 2e-6


def test_linear_rejects_bad_arguments(lib, L):
    rc = lib.dvae_linear(None, 1, 0, None, 1, 0, None, 1, 1, 1, 1, None, None, 0.0, 0, None)
    assert rc == -1 and b"null" in lib.dvae_last_error_string()
    x = torch.zeros(4, device="cuda")
    rc = lib.dvae_linear(L.ptr(x), 1, 0, L.ptr(x), 1, 0, L.ptr(x), 1, 0, 1, 1, None, None, 0.0, 0, None)
    assert rc == -1


# ---------------------------------------------------------------------------------------------
def _lstm_case(lib, L, T, B, I, H, D, with_len, with_h0, seed, dw_planes=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, B, I)).astype(np.float32)
    k = 1.0 / np.sqrt(H)
    W = [dict(w_ih=rng.uniform(-k, k, (4 * H, I)).astype(np.float32), w_hh=rng.uniform(-k, k, (4 * H, H)).astype(np.float32),
              b_ih=rng.uniform(-k, k, 4 * H).astype(np.float32), b_hh=rng.uniform(-k, k, 4 * H).astype(np.float32))
         for _ in range(D)]
    lengths = None
    if with_len:
        lengths = rng.integers(1, T + 1, B)
        lengths[0] = T
    h0 = c0 = None
    if with_h0:
        h0 = rng.standard_normal((D, B, H)).astype(np.float32) * 0.5
        c0 = rng.standard_normal((D, B, H)).astype(np.float32) * 0.5
    d_hs = rng.standard_normal((T, B, D * H)).astype(np.float32)
    d_hn = rng.standard_normal((D, B, H)).astype(np.float32)
    d_cn = rng.standard_normal((D, B, H)).astype(np.float32)
    # oracle (float64)
    want = []
    for d in range(D):
        w = {k_: v.astype(np.float64) for k_, v in W[d].items()}
        z = np.zeros((B, H))
        hs, hn, cn, cache = O.lstm_seq_fwd(x.astype(np.float64), w["w_ih"], w["w_hh"], w["b_ih"], w["b_hh"],
                                           h0[d].astype(np.float64) if with_h0 else z,
                                           c0[d].astype(np.float64) if with_h0 else z, lengths, reverse=(d == 1))
        g = O.lstm_seq_bwd(cache, d_hs=d_hs[:, :, d * H:(d + 1) * H].astype(np.float64), d_hn=d_hn[d].astype(np.float64),
                           d_cn=d_cn[d].astype(np.float64))
        want.append((hs, hn, cn, g))
    # device
    st = L.stream_ptr()
    xd = dev(x)
    Wd = [{k_: dev(v) for k_, v in W[d].items()} for d in range(D)]
    len_d = dev(lengths, torch.int64) if with_len else None
    h0d, c0d = (dev(h0), dev(c0)) if with_h0 else (None, None)
    hs = torch.full((T, B, D * H), 7.0, device="cuda")
    hn = torch.zeros(D, B, H, device="cuda")
    cn = torch.zeros(D, B, H, device="cuda")
    gates = torch.zeros(D, T, B, 4 * H, device="cuda")
    cs = torch.zeros(D, T, B, H, device="cuda")
    ws = torch.zeros(lib.dvae_lstm_state_ws_floats(B, H, D), device="cuda")
    pa = lambda key: L.ptr_array([Wd[d][key] for d in range(D)])
    L.check(lib.dvae_lstm_seq_fwd(L.ptr(xd), I, T, B, I, H, D, pa("w_ih"), pa("w_hh"), pa("b_ih"), pa("b_hh"),
                                  L.ptr(h0d), L.ptr(c0d), H, B * H, L.ptr(len_d), L.ptr(hs), D * H, L.ptr(hn), L.ptr(cn),
                                  H, B * H, L.ptr(gates), L.ptr(cs), L.ptr(ws), st), "lstm fwd")
    for d in range(D):
        assert rel(hs[:, :, d * H:(d + 1) * H], want[d][0]) < 5e-6
        assert rel(hn[d], want[d][1]) < 5e-6
        assert rel(cn[d], want[d][2]) < 5e-6
    # the same layer as two calls (input projection on its own, then the recurrence alone): bit-identical outputs
    hs2, hn2, cn2 = torch.full_like(hs, 5.0), torch.zeros_like(hn), torch.zeros_like(cn)
    gates2, cs2 = torch.zeros_like(gates), torch.zeros_like(cs)
    L.check(lib.dvae_lstm_input_proj(L.ptr(xd), I, T, B, I, H, D, pa("w_ih"), pa("b_ih"), pa("b_hh"), L.ptr(gates2), st), "lstm proj")
    L.check(lib.dvae_lstm_seq_fwd_ex(L.ptr(xd), I, T, B, I, H, D, pa("w_ih"), pa("w_hh"), pa("b_ih"), pa("b_hh"),
                                     L.ptr(h0d), L.ptr(c0d), H, B * H, L.ptr(len_d), L.ptr(hs2), D * H, L.ptr(hn2), L.ptr(cn2),
                                     H, B * H, L.ptr(gates2), L.ptr(cs2), L.ptr(ws), 1, st), "lstm fwd (recurrence only)")
    live = torch.ones(T, B, dtype=torch.bool, device="cuda") if not with_len else \
        (torch.arange(T, device="cuda")[:, None] < len_d[None, :])
    if H in (64, 128, 256) or not os.environ.get("DVAE_LSTM_IMPL", "") == "":
        assert torch.equal(hs2[live], hs[live]) and torch.equal(hn2, hn) and torch.equal(cn2, cn)
    else:      # large-H plane path: split-K partial sums meet in fp32 atomics, whose order varies from launch to launch
        assert rel(hs2[live], hs[live].cpu().numpy()) < 1e-6 and rel(hn2, hn.cpu().numpy()) < 1e-6 and rel(cn2, cn.cpu().numpy()) < 1e-6
    # backward
    G = [{k_: torch.full_like(v, 3.0) for k_, v in Wd[d].items()} for d in range(D)]
    ga = lambda key: L.ptr_array([G[d][key] for d in range(D)])
    d_x = torch.zeros(T, B, I, device="cuda")
    d_h0 = torch.zeros(D, B, H, device="cuda")
    d_c0 = torch.zeros(D, B, H, device="cuda")
    bwd_args = (L.ptr(xd), I, T, B, I, H, D, pa("w_ih"), pa("w_hh"), L.ptr(h0d), L.ptr(c0d), H, B * H,
                L.ptr(len_d), L.ptr(hs), D * H, L.ptr(gates), L.ptr(cs), L.ptr(dev(d_hs)), D * H,
                L.ptr(dev(d_hn)), L.ptr(dev(d_cn)), H, B * H, L.ptr(d_x), I, ga("w_ih"), ga("w_hh"),
                ga("b_ih"), ga("b_hh"), L.ptr(d_h0) if with_h0 else None,
                L.ptr(d_c0) if with_h0 else None, H, B * H, L.ptr(ws))
    if dw_planes:     # weight gradients from transposed operand planes of x, dG and hs (the caller set DVAE_DW_PLANES=1)
        pws = torch.full((lib.dvae_lstm_bwd_planes_ws_floats(T, B, I, H, D),), float("nan"), device="cuda")
        L.check(lib.dvae_lstm_seq_bwd_ex(*bwd_args, L.ptr(pws), st), "lstm bwd (planes)")
    else:
        L.check(lib.dvae_lstm_seq_bwd(*bwd_args, st), "lstm bwd")
    dx_want = sum(want[d][3]["dx"] for d in range(D))
    assert rel(d_x, dx_want) < 1e-4
    for d in range(D):
        g = want[d][3]
        assert rel(G[d]["w_ih"], g["dw_ih"]) < 1e-4
        assert rel(G[d]["w_hh"], g["dw_hh"]) < 1e-4
        assert rel(G[d]["b_ih"], g["db_ih"]) < 1e-4
        assert rel(G[d]["b_hh"], g["db_hh"]) < 1e-4
        if with_h0:
            assert rel(d_h0[d], g["dh0"]) < 1e-4
            assert rel(d_c0[d], g["dc0"]) < 1e-4


@pytest.mark.parametrize("T,B,I,H,D,with_len,with_h0", [
    (5, 3, 6, 8, 1, False, True),        # decoder-like: initial state given, all rows run all steps
    (7, 5, 10, 16, 2, True, False),      # encoder-like: bidirectional, ragged lengths
    (1, 4, 5, 4, 2, True, False),        # single step
    (9, 33, 12, 20, 1, True, False),     # rows / units not multiples of the tile sizes
    (22, 128, 64, 256, 2, True, False),  # cfg-2 hidden size
    (6, 40, 256, 64, 1, False, True),
    (4, 7, 7, 5, 2, True, False),        # widths not divisible by 4 (scalar load path)
    (11, 50, 24, 128, 2, True, False),   # persistent-cluster kernels: H=128, ragged, partial 16-row slice
    (21, 128, 256, 256, 1, False, True), # persistent-cluster kernels: cfg-2 decoder layer
    (30, 64, 32, 64, 1, True, False),
    (9, 50, 32, 256, 2, True, False),    # tcgen05 kernels: partial 16-row slice, ragged, both directions in one launch
    (7, 40, 48, 256, 1, False, True),    # tcgen05 kernels: initial state + d_h0/d_c0, partial slice
    (1, 16, 32, 256, 1, False, True),    # tcgen05 kernels: single step
    (2, 20, 32, 256, 2, True, False),
])
def test_lstm_seq(lib, L, T, B, I, H, D, with_len, with_h0):
    _lstm_case(lib, L, T, B, I, H, D, with_len, with_h0, seed=T * 31 + B)


@pytest.mark.parametrize("T,B,I,H,D,with_len,with_h0", [
    (6, 20, 64, 512, 2, True, False),     # large-H path (lstm_planes.cu): per-step plane GEMMs, bidirectional, ragged, partial row block
    (5, 128, 32, 1024, 1, False, True),   # cfg-4 hidden size, decoder-like: initial state, d_h0 / d_c0
    (9, 130, 48, 384, 2, True, False),    # H not a power of two, two row blocks (one partial)
    (1, 8, 32, 512, 1, False, True),      # single step
    (3, 16, 1024, 1024, 2, True, False),  # cfg-4 encoder layer-1-like input width
])
def test_lstm_seq_large_hidden_plane_path(lib, L, T, B, I, H, D, with_len, with_h0):
    _lstm_case(lib, L, T, B, I, H, D, with_len, with_h0, seed=T * 17 + B + H)


@pytest.mark.parametrize("T,B,I,H,D,with_len,with_h0", [
    (22, 128, 64, 256, 2, True, False),   # both directions, ragged; B % 32 == 0: dW_hh pairs dG_t with h_{t-1} by a k-block shift
    (9, 50, 32, 256, 2, True, False),     # B % 32 != 0: dW_ih from planes, dW_hh on the converter path
    (21, 128, 256, 256, 1, False, True),  # initial state: the h0 . dG_0 term is added after the plane GEMM
    (5, 128, 32, 1024, 1, False, True),   # cfg-4 hidden size (where the planes are the default)
    (6, 96, 200, 512, 2, True, False),    # I not a multiple of the tile sizes
    (1, 128, 32, 256, 1, False, True),    # single step: no h_{t-1} pairs at all
    (3, 16, 32, 256, 1, False, True),     # T*B < 128: falls back to the converter GEMMs
])
def test_lstm_seq_bwd_weight_gradient_planes(lib, L, T, B, I, H, D, with_len, with_h0, monkeypatch):
    monkeypatch.setenv("DVAE_DW_PLANES", "1")
    _lstm_case(lib, L, T, B, I, H, D, with_len, with_h0, seed=T * 13 + B + I, dw_planes=True)


@pytest.mark.parametrize("splits", ["1", "4"])
def test_lstm_seq_large_hidden_split_k_settings(lib, L, splits, monkeypatch):
    monkeypatch.setenv("DVAE_PLANES_SPLITS", splits)
    _lstm_case(lib, L, 4, 40, 32, 512, 2, True, False, seed=int(splits) + 5)


def test_lstm_step_kernels_still_cover_large_hidden(lib, L, monkeypatch):
    """DVAE_LSTM_IMPL=step: the exact fp32 per-step kernels, the plane path's A/B partner."""
    monkeypatch.setenv("DVAE_LSTM_IMPL", "step")
    _lstm_case(lib, L, 4, 20, 32, 512, 2, True, False, seed=3)


@pytest.mark.parametrize("groups,T,B,D,with_len,with_h0", [("32", 9, 50, 2, True, False), ("32", 7, 128, 1, False, True), ("32", 3, 33, 2, True, False),
                                                           ("2", 9, 50, 2, True, False), ("1", 6, 40, 1, False, True)])
def test_lstm_tc_row_group_layouts(lib, L, groups, T, B, D, with_len, with_h0, monkeypatch):
    """tcgen05 LSTM kernels, H = 256: one 32-row group per cluster (the layout a bidirectional layer at B = 128 runs), two
    16-row groups (round 1's), one 16-row group -- forced through DVAE_LSTM_GROUPS on shapes with partial row groups."""
    monkeypatch.setenv("DVAE_LSTM_GROUPS", groups)
    _lstm_case(lib, L, T, B, 48, 256, D, with_len, with_h0, seed=int(groups) * 7 + T + B)


@pytest.mark.parametrize("D,with_len,with_h0", [(2, True, False), (1, False, True)])
def test_lstm_simt_persistent_path_matches_oracle_too(lib, L, D, with_len, with_h0, monkeypatch):
    """DVAE_LSTM_IMPL=simt keeps the fp32 SIMT persistent-cluster kernels at H=256 (the tcgen05 kernels' A/B partner)."""
    monkeypatch.setenv("DVAE_LSTM_IMPL", "simt")
    _lstm_case(lib, L, 9, 48, 32, 256, D, with_len, with_h0, seed=77 + D)


@pytest.mark.parametrize("H,D,with_len,with_h0", [(256, 2, True, False), (64, 1, False, True)])
def test_lstm_step_path_matches_oracle_too(lib, L, H, D, with_len, with_h0, monkeypatch):
    """DVAE_LSTM_IMPL=step forces the general per-step kernels on shapes the persistent kernels take."""
    monkeypatch.setenv("DVAE_LSTM_IMPL", "step")
    _lstm_case(lib, L, 9, 48, 32, H, D, with_len, with_h0, seed=H + D)


# ---------------------------------------------------------------------------------------------
def _heads_oracle(ctx, Wc, bc, eps, dims, dsc_out, Wd, bd, labels, klw, Wz, bz, d_hid, B):
    """float64 forward + backward of the fused heads, written directly from vae/model.py:384-411."""
    out = {}
    Z = sum(dims)
    zcat, mus, lvs = [], [], []
    off = 0
    kls, dls, das, logits = [], [], [], []
    woff = boff = 0
    for s, zs in enumerate(dims):
        p = ctx @ Wc[2 * off:2 * off + 2 * zs].T + bc[2 * off:2 * off + 2 * zs]
        mu, lv = p[:, :zs], np.tanh(p[:, zs:])
        z = mu + eps[:, off:off + zs] * np.exp(lv)
        zcat.append(z); mus.append(mu); lvs.append(lv)
        kls.append(O.kl_divergence(mu, lv))
        if dsc_out[s] > 0:
            o = dsc_out[s]
            W = Wd[woff:woff + o * zs].reshape(o, zs)
            lg = z @ W.T + bd[boff:boff + o]
            logits.append(lg)
            woff += o * zs; boff += o
        off += zs
    out["z"], out["mu"], out["logvar"] = (np.concatenate(v, 1) for v in (zcat, mus, lvs))
    out["kl"] = np.array(kls)
    out["hid"] = np.tanh(out["z"] @ Wz.T + bz)
    out["logits"] = np.concatenate(logits, 1) if logits else np.zeros((B, 0))
    return out


@pytest.mark.parametrize("B,C,dims,dsc_out", [
    (6, 32, [1, 2, 4], [1, 1, 0]),
    (128, 1024, [1, 1, 62], [1, 1, 0]),
    (9, 24, [2, 3], [3, 0]),
    (5, 16, [4], [0]),
])
def test_latent_heads_forward(lib, L, B, C, dims, dsc_out):
    rng = np.random.default_rng(B + C)
    S, Z = len(dims), sum(dims)
    H2L = 4 * 12
    ctx = rng.standard_normal((B, C)).astype(np.float32) * 0.5
    Wc = rng.uniform(-1, 1, (2 * Z, C)).astype(np.float32) / np.sqrt(C)
    bc = rng.uniform(-0.1, 0.1, 2 * Z).astype(np.float32)
    eps = rng.standard_normal((B, Z)).astype(np.float32)
    n_w = sum(o * z for o, z in zip(dsc_out, dims))
    OD = sum(dsc_out)
    Wd = rng.uniform(-1, 1, max(n_w, 1)).astype(np.float32)
    bd = rng.uniform(-1, 1, max(OD, 1)).astype(np.float32)
    Wz = rng.uniform(-1, 1, (H2L, Z)).astype(np.float32) / np.sqrt(Z)
    bz = rng.uniform(-0.1, 0.1, H2L).astype(np.float32)
    nd = sum(1 for o in dsc_out if o > 0)
    labels = np.zeros((max(nd, 1), B), np.float32)
    i = 0
    for o in dsc_out:
        if o > 0:
            labels[i] = rng.integers(0, max(o, 2), B)
            i += 1
    klw = rng.uniform(0, 1, S).astype(np.float32)
    f64 = lambda a: a.astype(np.float64)
    want = _heads_oracle(f64(ctx), f64(Wc), f64(bc), f64(eps), dims, dsc_out, f64(Wd), f64(bd), labels, klw, f64(Wz), f64(bz), None, B)
    z, mu, lv = (torch.zeros(B, Z, device="cuda") for _ in range(3))
    hid = torch.zeros(B, H2L, device="cuda")
    lg = torch.zeros(B, max(OD, 1), device="cuda")
    sc = torch.zeros(L.HEADS_NSCALARS, device="cuda")
    ws = torch.zeros(lib.dvae_heads_ws_floats(B, S), device="cuda")
    L.check(lib.dvae_latent_heads_fwd(L.ptr(dev(ctx)), B, C, S, L.int_array(dims), L.int_array(dsc_out), L.ptr(dev(Wc)),
                                      L.ptr(dev(bc)), L.ptr(dev(eps)), L.ptr(dev(Wd)), L.ptr(dev(bd)), L.ptr(dev(labels)),
                                      L.ptr(dev(klw)), L.ptr(dev(Wz)), L.ptr(dev(bz)), H2L, L.ptr(z), L.ptr(mu), L.ptr(lv),
                                      L.ptr(hid), L.ptr(lg), L.ptr(sc), L.ptr(ws), L.stream_ptr()), "heads fwd")
    assert rel(z, want["z"]) < 5e-6 and rel(mu, want["mu"]) < 5e-6 and rel(lv, want["logvar"]) < 5e-6
    assert rel(hid, want["hid"]) < 5e-6
    if OD:
        assert rel(lg[:, :OD], want["logits"]) < 5e-6
    sc = sc.cpu().numpy()
    assert np.abs(sc[3:3 + S] - want["kl"]).max() < 1e-5 * max(1.0, np.abs(want["kl"]).max())
    assert abs(sc[0] - float((klw * want["kl"]).sum())) < 1e-5 * max(1.0, abs(float((klw * want["kl"]).sum())))
    assert abs(sc[1] - want["kl"].sum()) < 1e-5 * max(1.0, want["kl"].sum())
    # discriminator losses
    off = 0
    i = 0
    tot = 0.0
    for s, o in enumerate(dsc_out):
        if o == 0:
            continue
        l = want["logits"][:, off:off + o]
        y = labels[i]
        if o == 1:
            x = l[:, 0]
            loss = (np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x)))).mean()
            acc = ((x > 0) == (y > 0.5)).mean()
        else:
            m = l.max(1)
            loss = (m + np.log(np.exp(l - m[:, None]).sum(1)) - l[np.arange(B), y.astype(int)]).mean()
            acc = (l.argmax(1) == y.astype(int)).mean()
        assert abs(sc[3 + S + s] - loss) < 1e-5 and abs(sc[3 + 2 * S + s] - acc) < 1e-6
        tot += loss
        off += o
        i += 1
    assert abs(sc[2] - tot) < 1e-5


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T1,B,H,V", [(4, 3, 8, 23), (7, 5, 16, 37), (19, 128, 256, 10000), (3, 130, 20, 300), (5, 9, 12, 1000), (6, 50, 64, 3001), (21, 128, 256, 10000),
                                      (5, 60, 512, 2050), (3, 100, 320, 1100)])     # H > 256: pre-split planes without stationary A
def test_vocab_ce(lib, L, T1, B, H, V):
    rng = np.random.default_rng(T1 + B + V)
    N, T = T1 * B, T1 + 1
    h = rng.standard_normal((T1, B, H)).astype(np.float32) * 0.5
    W = rng.uniform(-1, 1, (V, H)).astype(np.float32) / np.sqrt(H) * 3
    b = rng.uniform(-0.1, 0.1, V).astype(np.float32)
    lengths = rng.integers(1, T + 1, B)
    lengths[0] = T
    targets = rng.integers(0, V, (B, T))
    sos = 2
    targets[:, 0] = sos
    targets[1 % B, 0] = 5       # a row whose first target is not <SOS>
    logits = np.zeros((B, T, V))
    logits[:, 0, sos] = 1.0
    logits[:, 1:, :] = (h.astype(np.float64) @ W.astype(np.float64).T + b).transpose(1, 0, 2)
    want_loss, cache = O.seq_ce_fwd(logits, targets, lengths)
    dl = O.seq_ce_bwd(cache)[:, 1:, :].transpose(1, 0, 2).reshape(N, V)
    hd, Wd, bd = dev(h), dev(W), dev(b)
    td, ld = dev(targets, torch.int64), dev(lengths, torch.int64)
    lse, nll = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    am = torch.zeros(N, device="cuda", dtype=torch.int32)
    loss = torch.zeros(1, device="cuda")
    ws = torch.zeros(lib.dvae_vocab_ce_ws_floats(N, V, H), device="cuda")
    st = L.stream_ptr()
    L.check(lib.dvae_vocab_ce_fwd(L.ptr(hd), H, T1, B, H, V, L.ptr(Wd), L.ptr(bd), L.ptr(td), T, L.ptr(ld), sos,
                                  L.ptr(lse), L.ptr(nll), L.ptr(am), L.ptr(loss), L.ptr(ws), st), "ce fwd")
    assert abs(loss.item() - want_loss) < 1e-5 * abs(want_loss)
    assert rel(nll.view(T1, B), cache["nll"][:, 1:].T) < 1e-5
    ref_am = logits[:, 1:, :].argmax(-1).T.reshape(N)
    got_am = am.cpu().numpy()
    if not np.array_equal(got_am, ref_am):      # only exact fp32 near-ties may differ
        bad = np.nonzero(got_am != ref_am)[0]
        flat = logits[:, 1:, :].transpose(1, 0, 2).reshape(N, V)
        gap = np.abs(flat[bad, got_am[bad]] - flat[bad, ref_am[bad]])
        assert gap.max() < 1e-6, f"argmax differs on {len(bad)} rows with gap {gap.max()}"
    # backward
    d_h = torch.zeros(N, H, device="cuda")
    d_w = torch.full((V, H), 9.0, device="cuda")
    d_b = torch.full((V,), 9.0, device="cuda")
    wsb = torch.zeros(lib.dvae_vocab_ce_bwd_ws_floats(N, V, H), device="cuda")
    gs = torch.tensor([1.0], device="cuda")
    L.check(lib.dvae_vocab_ce_bwd(L.ptr(hd), H, T1, B, H, V, L.ptr(Wd), L.ptr(bd), L.ptr(td), T, L.ptr(ld), L.ptr(lse),
                                  L.ptr(gs), L.ptr(d_h), H, L.ptr(d_w), L.ptr(d_b), None, L.ptr(wsb), st), "ce bwd")
    hf = h.reshape(N, H).astype(np.float64)
    assert rel(d_h, dl @ W.astype(np.float64)) < 1e-4
    assert rel(d_w, dl.T @ hf) < 1e-4
    assert rel(d_b, dl.sum(0)) < 1e-4
    # W planes prepared ahead of the forward call (dvae_vocab_split_w + flags bit 0): same loss, lse and arg-max
    ws3 = torch.zeros_like(ws)
    lse3, nll3 = torch.zeros_like(lse), torch.zeros_like(lse)
    am3, loss3 = torch.zeros(N, device="cuda", dtype=torch.int32), torch.zeros(1, device="cuda")
    L.check(lib.dvae_vocab_split_w(L.ptr(Wd), N, V, H, L.ptr(ws3), st), "split w")
    L.check(lib.dvae_vocab_ce_fwd_ex(L.ptr(hd), H, T1, B, H, V, L.ptr(Wd), L.ptr(bd), L.ptr(td), T, L.ptr(ld), sos, L.ptr(lse3),
                                     L.ptr(nll3), L.ptr(am3), L.ptr(loss3), L.ptr(ws3), 1, st), "ce fwd_ex")
    assert torch.equal(lse3, lse) and torch.equal(loss3, loss) and torch.equal(am3, am)
    # the same call reusing the forward call's operand planes (fwd_ws) must give the same gradients (split-K atomics: not bitwise)
    d_h2, d_w2, d_b2 = torch.full_like(d_h, 3.0), torch.full_like(d_w, 3.0), torch.full_like(d_b, 3.0)
    L.check(lib.dvae_vocab_ce_bwd(L.ptr(hd), H, T1, B, H, V, L.ptr(Wd), L.ptr(bd), L.ptr(td), T, L.ptr(ld), L.ptr(lse),
                                  L.ptr(gs), L.ptr(d_h2), H, L.ptr(d_w2), L.ptr(d_b2), L.ptr(ws), L.ptr(wsb), st), "ce bwd (fwd_ws)")
    assert rel(d_h2, d_h.cpu().numpy().astype(np.float64)) < 1e-6 and rel(d_w2, d_w.cpu().numpy().astype(np.float64)) < 1e-6
    assert rel(d_b2, d_b.cpu().numpy().astype(np.float64)) < 1e-6


def test_vocab_ce_cta_pair_multicast(lib, L, monkeypatch):
    """DVAE_TC16_MCAST=1: CTA pairs fetch half of every B stage each and multicast it (odd row-block count: padding CTA)."""
    monkeypatch.setenv("DVAE_TC16_MCAST", "1")
    test_vocab_ce(lib, L, 6, 50, 64, 3001)         # 3 row blocks -> 4 CTAs along x
    test_vocab_ce(lib, L, 19, 128, 256, 10000)


def test_vocab_ce_simt_path_matches_oracle_too(lib, L, monkeypatch):
    """DVAE_GEMM_IMPL=simt forces the fp32 SIMT kernels on a shape the tensor-core path takes."""
    monkeypatch.setenv("DVAE_GEMM_IMPL", "simt")
    test_vocab_ce(lib, L, 19, 128, 256, 10000)


# ---------------------------------------------------------------------------------------------
def test_clip_adam(lib, L):
    rng = np.random.default_rng(5)
    n = 100003
    p = rng.standard_normal(n).astype(np.float32)
    g = (rng.standard_normal(n) * 0.1).astype(np.float32)
    m = (rng.standard_normal(n) * 0.01).astype(np.float32)
    v = (rng.uniform(0, 1, n) * 1e-3).astype(np.float32)
    pd, gd, md, vd = dev(p), dev(g), dev(m), dev(v)
    ss = torch.zeros(1, device="cuda")
    ws = torch.zeros(2048, device="cuda")
    hyper = dev(np.array([3e-4, 0.9, 0.999, 1e-8, 7.0], np.float32))
    st = L.stream_ptr()
    for _ in range(2):     # the reduction re-arms its own counter
        L.check(lib.dvae_grad_sumsq(L.ptr(gd), n, L.ptr(ss), L.ptr(ws), st), "sumsq")
        assert abs(ss.item() - float((g.astype(np.float64) ** 2).sum())) < 1e-5 * float((g.astype(np.float64) ** 2).sum())
    L.check(lib.dvae_clip_adam(L.ptr(pd), L.ptr(gd), L.ptr(md), L.ptr(vd), n, L.ptr(ss), 5.0, 1.0, L.ptr(hyper), 1, st), "adam")
    P, G, M, V = {"w": p.astype(np.float64)}, {"w": g.astype(np.float64)}, {"w": m.astype(np.float64)}, {"w": v.astype(np.float64)}
    norm = O.clip_and_adam(P, G, M, V, step=7, lr=3e-4)
    assert norm > 5.0      # the clip is active in this case
    assert rel(md, M["w"]) < 1e-5 and rel(vd, V["w"]) < 1e-5
    assert np.abs(pd.cpu().numpy() - P["w"]).max() < 1e-6
    assert float(gd.abs().max()) == 0.0


def test_embedding_dropout_roundtrip(lib, L):
    rng = np.random.default_rng(1)
    V, E, T, B = 50, 20, 6, 9
    emb = rng.standard_normal((V, E)).astype(np.float32)
    tok = rng.integers(0, V, (B, T + 3))
    embd, tokd = dev(emb), dev(tok, torch.int64)
    seed = torch.tensor([1234567], device="cuda", dtype=torch.int64)
    x0 = torch.zeros(T, B, E, device="cuda")
    st = L.stream_ptr()
    L.check(lib.dvae_embedding_fwd(L.ptr(embd), E, L.ptr(tokd), T + 3, 1, T, B, 0.0, None, 1, 2, 0, L.ptr(x0), st), "emb")
    want = emb[tok[:, :T].T]
    want[0] = emb[2]
    assert np.array_equal(x0.cpu().numpy(), want)
    x1 = torch.zeros_like(x0)
    x2 = torch.zeros_like(x0)
    for x in (x1, x2):
        L.check(lib.dvae_embedding_fwd(L.ptr(embd), E, L.ptr(tokd), T + 3, 1, T, B, 0.5, L.ptr(seed), 1, 2, 0, L.ptr(x), st), "emb")
    assert torch.equal(x1, x2)                      # same seed -> same mask
    mask = (x1 != 0)
    assert torch.allclose(x1[mask], 2.0 * x0[mask])
    keep = mask.float().mean().item()
    assert 0.4 < keep < 0.6
    # dense dropout uses the same counter layout -> same mask for the same (seed, salt, shape)
    y = torch.zeros_like(x0)
    L.check(lib.dvae_dropout(L.ptr(x0), E, T * B, E, 0.5, L.ptr(seed), 1, L.ptr(y), E, 0, st), "dropout")
    # a single-step call (t0 = 2, T = 1) reproduces that slice of the full-sequence call bit for bit
    x3 = torch.zeros_like(x0)
    L.check(lib.dvae_embedding_fwd(L.ptr(embd), E, L.ptr(tokd), T + 3, 1, 1, B, 0.5, L.ptr(seed), 1, 2, 2, L.ptr(x3), st), "emb")
    assert torch.equal(x3[2], x1[2]) and float(x3[3:].abs().max()) == 0
    assert torch.equal(y, x1)
    # backward scatter-add with the same mask
    d_x = dev(rng.standard_normal((T, B, E)).astype(np.float32))
    d_emb = torch.zeros(V, E, device="cuda")
    L.check(lib.dvae_embedding_bwd(L.ptr(d_x), E, L.ptr(tokd), T + 3, 1, T, B, 0.5, L.ptr(seed), 1, 2, 0, L.ptr(d_emb), st), "emb bwd")
    want = np.zeros((V, E))
    toks = tok[:, :T].T.copy()
    toks[0] = 2
    np.add.at(want, toks.reshape(-1), (d_x.cpu().numpy() * mask.cpu().numpy() * 2.0).reshape(-1, E))
    assert rel(d_emb, want) < 1e-5


def test_randn_statistics(lib, L):
    n = 1 << 20
    out = torch.zeros(n, device="cuda")
    seed = torch.tensor([99], device="cuda", dtype=torch.int64)
    L.check(lib.dvae_randn(L.ptr(out), n, L.ptr(seed), 64, L.stream_ptr()), "randn")
    assert abs(out.mean().item()) < 5e-3 and abs(out.std().item() - 1.0) < 5e-3
    assert abs((out ** 4).mean().item() - 3.0) < 0.05
