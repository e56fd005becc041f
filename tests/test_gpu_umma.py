"""tcgen05 building blocks in isolation: one UMMA GEMM through the sampler's operand paths (epilogue-style swizzled
bf16 store for A, TMA SWIZZLE_128B tiles for W, tcgen05.ld 32x32b read-back) against a torch fp32 matmul of the
same bf16-rounded operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mt,n,k", [(1, 64, 64), (2, 128, 128), (1, 256, 256), (3, 64, 512), (1, 512, 128)])
def test_umma_gemm(mt, n, k):
    import ddqst_b200 as dq
    lib = dq._lib.load()
    g = torch.Generator().manual_seed(n * 7 + k)
    a = torch.randn(128 * mt, k, generator=g)
    w = torch.randn(n, k, generator=g)
    a_d = a.cuda()
    w_bf = w.to(torch.bfloat16).cuda().contiguous()
    c = torch.full((128 * mt, n), float("nan"), device="cuda")
    dq._lib.check(lib.ddqst_selftest_umma(dq._lib.ptr(a_d), dq._lib.ptr(w_bf), mt, n, k, dq._lib.ptr(c), dq._lib.stream_ptr()))
    assert lib.ddqst_debug_tc_status() == 0, "tcgen05 pipeline timed out"
    want = a.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t()
    err = (c.cpu() - want).abs().max().item()
    assert err < 2e-3 * np.sqrt(k), err


@pytest.mark.parametrize("mp,n,k", [(1, 64, 64), (1, 256, 128), (2, 128, 512), (1, 16, 256), (1, 256, 256)])
def test_umma_gemm_cta_pair(mp, n, k):
    """cta_group::2: one MMA spans a CTA pair (M=256), each CTA stages half of W."""
    import ddqst_b200 as dq
    lib = dq._lib.load()
    g = torch.Generator().manual_seed(n * 11 + k)
    a = torch.randn(256 * mp, k, generator=g)
    w = torch.randn(n, k, generator=g)
    a_d = a.cuda()
    w_bf = w.to(torch.bfloat16).cuda().contiguous()
    c = torch.full((256 * mp, n), float("nan"), device="cuda")
    dq._lib.check(lib.ddqst_selftest_umma2(dq._lib.ptr(a_d), dq._lib.ptr(w_bf), mp, n, k, dq._lib.ptr(c), dq._lib.stream_ptr()))
    assert lib.ddqst_debug_tc_status() == 0, "tcgen05 pipeline timed out"
    want = a.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t()
    err = (c.cpu() - want).abs().max().item()
    assert err < 2e-3 * np.sqrt(k), err
