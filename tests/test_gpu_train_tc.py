"""Tensor-core training step (ddqst_train_forward_backward_tc, train_tc.cu) against (a) a torch fp32 matmul for the GEMM
kernel alone in all four operand storage orders, and (b) the fp32 CUDA-core training step, which itself is pinned to the
reference's losses / weights at 1e-5 (test_gpu_parity.py::test_train_steps_match_reference).

Tolerance: north_star allows 1e-2 relative error for bf16/tf32 denoiser arithmetic; gradients are compared per
parameter tensor in relative L2 norm."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dq():
    import ddqst_b200
    assert torch.cuda.is_available()
    return ddqst_b200


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k,batch", [(128, 64, 64, 1), (256, 128, 192, 2), (200, 96, 104, 1), (24, 512, 1000, 1),
                                         (512, 32, 520, 3), (1024, 512, 512, 1)])
def test_gemm_tc_all_majors(dq, a_mn, b_mn, m, n, k, batch):
    """C = A . B^T with either operand stored K-major ([mn,k]) or MN-major ([k,mn]); ragged M/N/K tails ride on
    TMA's zero fill."""
    lib = dq._lib.load()
    g = torch.Generator().manual_seed(m * 3 + n * 5 + k * 7 + a_mn * 11 + b_mn * 13)
    a = torch.randn(batch, m, k, generator=g).to(torch.bfloat16)
    b = torch.randn(batch, n, k, generator=g).to(torch.bfloat16)
    a_dev = (a.transpose(1, 2).contiguous() if a_mn else a.contiguous()).cuda()
    b_dev = (b.transpose(1, 2).contiguous() if b_mn else b.contiguous()).cuda()
    c = torch.full((batch, m, n), float("nan"), device="cuda")
    dq._lib.check(lib.ddqst_selftest_gemm_tc(dq._lib.ptr(a_dev), dq._lib.ptr(b_dev), a_mn, b_mn, m, n, k, batch,
                                             dq._lib.ptr(c), dq._lib.stream_ptr()))
    assert lib.ddqst_debug_tc_status() == 0, "tcgen05 pipeline timed out"
    want = torch.bmm(a.float(), b.float().transpose(1, 2))
    err = (c.cpu() - want).abs().max().item()
    assert err < 1e-3 * np.sqrt(k), err


def _grads(dq, m, diff, x0p, xtp, t32, b32, tc):
    lib = dq._lib.load()
    B = x0p.shape[0]
    grads = torch.full_like(m.flat_params, float("nan"))
    loss = torch.zeros(1, device="cuda")
    prec = dq._lib.PRECISION_BF16 if tc else dq._lib.PRECISION_FP32
    nbytes = lib.ddqst_workspace_bytes(dq._lib.OP_TRAIN, C.byref(m.dims), B, prec)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    P = dq._lib.ptr
    if tc:
        dq._lib.check(lib.ddqst_train_forward_backward_tc(C.byref(m.dims), P(m.flat_params), P(m.bf16_shadow()), P(xtp), P(x0p),
                                                          P(t32), P(b32), B, 1.0, P(grads), P(loss), P(ws), ws.numel(),
                                                          dq._lib.stream_ptr()))
        assert lib.ddqst_debug_tc_status() == 0, "tcgen05 pipeline timed out"
    else:
        dq._lib.check(lib.ddqst_train_forward_backward(C.byref(m.dims), P(m.flat_params), P(xtp), P(x0p), P(t32), P(b32), B, 1.0,
                                                       P(grads), P(loss), P(ws), ws.numel(), dq._lib.stream_ptr()))
    torch.cuda.synchronize()
    return loss.item(), grads


def _compare(dq, m, B, T, tol_loss, tol_grad, seed=0):
    N = m.num_qubits
    g = torch.Generator().manual_seed(seed)
    x0p = torch.randint(0, 1 << N, (B,), generator=g).to(torch.uint16).cuda()
    xtp = torch.randint(0, 1 << N, (B,), generator=g).to(torch.uint16).cuda()
    t32 = torch.randint(1, T + 1, (B,), generator=g).to(torch.int32).cuda()
    b32 = torch.randint(0, m.dims.num_bases, (B,), generator=g).to(torch.int32).cuda()
    diff = None
    l32, g32 = _grads(dq, m, diff, x0p, xtp, t32, b32, tc=False)
    ltc, gtc = _grads(dq, m, diff, x0p, xtp, t32, b32, tc=True)
    assert np.isfinite(ltc) and abs(ltc - l32) <= tol_loss * abs(l32), (ltc, l32)
    assert torch.isfinite(gtc).all()
    worst = {}
    names = [n for n, _ in m.named_parameters()]
    for name, a, b in zip(names, m.views_of(g32), m.views_of(gtc)):
        denom = a.norm().item()
        rel = (a - b).norm().item() / max(denom, 1e-12)
        worst[name] = rel
    bad = {k: v for k, v in worst.items() if v > tol_grad}
    assert not bad, f"gradient mismatch (relative L2): {bad}"
    return worst


@pytest.mark.parametrize("tag", ["B", "A"])
def test_train_tc_matches_fp32_small(dq, tag):
    """The reference-fixture models (N=3, E=16, H=64, L=2; B = RQC x_emb front end, A = SS Linear(N,H) front end):
    ragged batch, tiny K/N dims -> every TMA tail path."""
    z = load_golden(f"model_{tag}_small.npz")
    N, NB, T, E, H, L = (int(v) for v in z["dims"])
    m = dq.ConditionalD3PM(N, NB, T, E, H, L, variant=tag)
    m.load_state_dict(golden_state_dict(z, "sd."))
    m = m.cuda()
    _compare(dq, m, 300, T, tol_loss=5e-3, tol_grad=3e-2)


def test_train_tc_matches_fp32_c4(dq):
    """Config C4's model (N=8, E=128, H=512, L=4, 6561 bases), batch 1024 (RQC/config.py:15)."""
    torch.manual_seed(0)
    m = dq.ConditionalD3PM(8, 6561, 100, 128, 512, 4).cuda()
    _compare(dq, m, 1024, 100, tol_loss=5e-3, tol_grad=3e-2)


def test_train_tc_matches_fp32_wide_tiles(dq):
    """Batch large enough that the fused-epilogue GEMMs take the 128x128 tile path (4 accumulator chunks per row, 6-stage
    ring) and the grouped weight-gradient launch spans several waves; ragged last row tile (6200 = 48*128 + 56)."""
    torch.manual_seed(5)
    m = dq.ConditionalD3PM(4, 81, 50, 32, 256, 2).cuda()
    _compare(dq, m, 6200, 50, tol_loss=5e-3, tol_grad=3e-2, seed=2)


def test_train_tc_steps_track_fp32(dq):
    """30 optimiser steps on the same data and noise stream: the tensor-core run's loss curve follows the fp32 run's."""
    torch.manual_seed(1)
    N, T = 4, 50
    mk = lambda: dq.ConditionalD3PM(N, 81, T, 32, 128, 2).cuda()
    m32, mtc = mk(), mk()
    mtc.load_state_dict(m32.state_dict())
    g = torch.Generator().manual_seed(3)
    probs = torch.rand(1 << N, generator=g) ** 3
    x0 = torch.multinomial(probs, 2048, replacement=True, generator=g).to(torch.uint16).cuda()
    basis = torch.randint(0, 81, (2048,), generator=g).cuda()
    d32 = dq.DiscreteDiffusion(m32, T, "cuda", seed=5, precision="fp32")
    dtc = dq.DiscreteDiffusion(mtc, T, "cuda", seed=5, precision="bf16")
    assert dtc.train_precision() == "bf16" and d32.train_precision() == "fp32"
    o32, otc = dq.NativeAdam(m32, lr=1e-3), dq.NativeAdam(mtc, lr=1e-3)
    l32 = [d32.train_step(x0, basis, o32).item() for _ in range(30)]
    ltc = [dtc.train_step(x0, basis, otc).item() for _ in range(30)]
    assert dq._lib.load().ddqst_debug_tc_status() == 0
    assert l32[-1] < l32[0] and ltc[-1] < ltc[0]
    assert np.abs(np.array(l32) - np.array(ltc)).max() < 2e-2 * max(l32), (l32, ltc)
    rel = (m32.flat_params - mtc.flat_params).norm().item() / m32.flat_params.norm().item()
    assert rel < 2e-2, rel
    # the bf16 shadow the Adam kernel maintains equals a fresh cast of the parameters
    assert torch.equal(mtc.bf16_shadow(), mtc.flat_params.to(torch.bfloat16))


def test_train_tc_variant_a_adamw_tracks_reference(dq):
    """SS phase: variant A model, linear schedule, AdamW(lr 1e-4, wd 0.01) (SS/main.py:77) on the reference fixture."""
    z = load_golden("model_A_small.npz")
    N, NB, T, E, H, L = (int(v) for v in z["dims"])
    m = dq.ConditionalD3PM(N, NB, T, E, H, L, variant="A")
    m.load_state_dict(golden_state_dict(z, "sd."))
    m = m.cuda()
    diff = dq.DiscreteDiffusion(m, T, "cuda", schedule="linear", seed=int(z["train_seed"][0]))
    assert diff.train_precision() == "bf16"
    opt = dq.NativeAdam(m, lr=1e-4, weight_decay=0.01, decoupled=True)
    x0, b0 = torch.from_numpy(z["train_x0"]).cuda(), torch.from_numpy(z["train_basis"]).cuda()
    losses = [diff.train_step(x0, b0, opt).item() for _ in range(3)]
    assert dq._lib.load().ddqst_debug_tc_status() == 0
    assert np.allclose(losses, z["train_losses"], rtol=1e-2), (losses, z["train_losses"])
    want = golden_state_dict(z, "trained.")
    for k, v in want.items():
        assert (m.state_dict()[k].cpu() - v).abs().max().item() < 2.5e-4, k      # AdamW moves each weight by <= lr per step


def test_train_graph_replay_matches_eager(dq):
    """A captured step replayed k times == k eager tensor-core steps (same device-side step counter / noise stream)."""
    torch.manual_seed(2)
    N, T, B = 4, 50, 512
    mk = lambda: dq.ConditionalD3PM(N, 81, T, 32, 128, 2).cuda()
    me, mg = mk(), mk()
    mg.load_state_dict(me.state_dict())
    g = torch.Generator().manual_seed(4)
    x0 = torch.randint(0, 1 << N, (B,), generator=g).to(torch.uint16).cuda()
    basis = torch.randint(0, 81, (B,), generator=g).to(torch.int32).cuda()
    de = dq.DiscreteDiffusion(me, T, "cuda", seed=9, precision="bf16")
    dg = dq.DiscreteDiffusion(mg, T, "cuda", seed=9, precision="bf16")
    oe, og = dq.NativeAdam(me, lr=1e-3), dq.NativeAdam(mg, lr=1e-3)
    tg = dg.make_train_graph(x0, basis, og)          # runs one warm-up step, then captures
    eager = [de.train_step(x0, basis, oe).item() for _ in range(4)]
    replay = [tg.replay().item() for _ in range(3)]
    assert dq._lib.load().ddqst_debug_tc_status() == 0
    # the embedding scatter uses fp32 atomics, so runs agree to rounding, not bitwise
    assert np.allclose(eager[1:], replay, rtol=2e-3, atol=2e-4), (eager, replay)
    assert int(og.step_dev[0].item()) == 4 and og.step_count == 4
    rel = (me.flat_params - mg.flat_params).norm().item() / me.flat_params.norm().item()
    assert rel < 1e-3, rel


@pytest.mark.parametrize("dims,B", [((8, 6561, 100, 128, 512, 4), 1024), ((8, 6561, 100, 128, 512, 4), 4200), ((4, 81, 50, 32, 256, 2), 6200), ((4, 81, 50, 32, 128, 2), 300),
                                    ((10, 59049, 100, 128, 512, 4), 700), ((3, 27, 100, 64, 512, 4), 256)])
def test_fused_forward_backward_kernel(dq, dims, B):
    """train_fused_kernel<H> (one persistent cta_group::2 launch for forward + data gradients, csrc/train_fused.cuh) forced on,
    against the fp32 CUDA-core step (itself pinned to the reference at 1e-5) and against the per-layer tensor-core path:
    C4 (also at 4200 rows: the 128 x 256 weight-gradient tiles) and C5 architectures, a ragged batch with an odd tile count (6200 = 48 tiles + 56 rows), H = 128 / 256 / 512, and the
    SS variant (E = 64).  Same bars as the per-layer path: loss 5e-3, every gradient tensor 3e-2 relative L2."""
    lib = dq._lib.load()
    variant = "A" if dims[3] == 64 else "B"
    torch.manual_seed(0)
    m = dq.ConditionalD3PM(*dims, variant=variant).cuda()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.02 * torch.randn_like(p))
    m.native_version += 1
    try:
        dq._lib.check(lib.ddqst_debug_train_path(1))
        worst_fused = _compare(dq, m, B, dims[2], tol_loss=5e-3, tol_grad=3e-2, seed=1)
        dq._lib.check(lib.ddqst_debug_train_path(0))
        worst_layer = _compare(dq, m, B, dims[2], tol_loss=5e-3, tol_grad=3e-2, seed=1)
    finally:
        dq._lib.check(lib.ddqst_debug_train_path(-1))
    # the fused pass keeps more intermediates in bf16 (z1, s, gamma|beta) than the per-layer path: its error may be larger, not by much
    assert max(worst_fused.values()) <= max(2.5 * max(worst_layer.values()), 1.5e-2), (worst_fused, worst_layer)
