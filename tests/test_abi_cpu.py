"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/ddqst.h declares; the host
logic of the Python mirror (flat parameters, state_dict contract, schedules, slot tables, sharding) is right."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_state_dict, load_golden
from oracle import ddqst_oracle as orc
from oracle import ref_harness as rh

import ddqst_b200 as dq


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ddqst.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddqst_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(dq._lib.library_path())
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ddqst.h but not exported"
    assert set(dq._lib.SIGNATURES) == set(names), set(dq._lib.SIGNATURES) ^ set(names)
    assert dq._lib.load().ddqst_version() >= 100


def test_layout_queries_work_without_a_gpu():
    lib = dq._lib.load()
    d = dq._lib.Dims(8, 6561, 100, 128, 512, 4, 1)
    offs = (ctypes.c_int64 * (5 + 24 + 2))()
    total = lib.ddqst_param_count(ctypes.byref(d), offs)
    assert total >= 4_539_920 and total < 4_539_920 + 4 * 31        # SURVEY 8a row M1 (+ alignment padding)
    assert list(offs) == sorted(offs) and offs[0] == 0
    assert lib.ddqst_pack_bytes(ctypes.byref(d)) > 6561 * 4 * 1024 * 4
    bad = dq._lib.Dims(8, 6561, 100, 128, 512, 40, 1)
    assert lib.ddqst_param_count(ctypes.byref(bad), None) == -1
    assert b"num_blocks" in lib.ddqst_last_error()


def test_compute_calls_fail_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = dq.ConditionalD3PM(2, 9, 10, 8, 64, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 2, dtype=torch.long), torch.ones(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        dq.DiscreteDiffusion(m, 10, "cpu").p_sample(4, 0, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        dq.linear_inversion(np.zeros((9, 4), np.int32), 2)
    lib = dq._lib.load()
    assert lib.ddqst_histogram(None, 1, 16, 3, None, None) == -2          # DDQST_EUNSUPPORTED_ARCH
    assert b"no CPU fallback" in lib.ddqst_last_error() or b"not supported" in lib.ddqst_last_error()


@pytest.mark.parametrize("tag", ["A", "B"])
def test_state_dict_contract(tag):
    z = load_golden(f"model_{tag}_small.npz")
    N, NB, T, E, H, L = (int(v) for v in z["dims"])
    m = dq.ConditionalD3PM(N, NB, T, E, H, L, variant=tag)
    sd = golden_state_dict(z)
    assert list(m.state_dict().keys()) == list(sd.keys()) or set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)                                          # strict: reference checkpoints load unchanged
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]) and v.shape == sd[k].shape
    # every parameter is a view into the single flat buffer, also after .to()/.float()
    m = m.float()
    base = m.flat_params.untyped_storage().data_ptr()
    assert all(p.untyped_storage().data_ptr() == base for p in m.parameters())
    assert sum(p.numel() for p in m.parameters()) <= m.flat_params.numel()
    v0 = tuple(p._version for p in m.parameters())
    before = m.flat_params.clone()
    with torch.no_grad():
        m.output_head.bias.add_(1.0)
    assert tuple(p._version for p in m.parameters()) != v0          # packed state will be rebuilt
    assert not torch.equal(before, m.flat_params)                   # ... and the write landed in the flat buffer


@pytest.mark.skipif(not rh.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("tag,phase", [("A", "SS"), ("B", "RQC")])
def test_default_init_equals_reference(tag, phase):
    ref = rh.load_phase(phase, names=("model",))["model"]
    torch.manual_seed(123)
    a = ref.ConditionalD3PM(3, 27, 20, 16, 64, 2)
    torch.manual_seed(123)
    b = dq.ConditionalD3PM(3, 27, 20, 16, 64, 2, variant=tag)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


def test_schedules_equal_oracle():
    for T in (10, 100):
        b, q = dq.cosine_schedule(T)
        ob, oq = orc.cosine_schedule(T)
        assert torch.equal(b, ob) and torch.equal(q, oq)
        b, q = dq.linear_schedule(T)
        ob, oq = orc.linear_schedule(T)
        assert torch.equal(b, ob) and torch.equal(q, oq)


def test_first_compatible_slot_table():
    from ddqst_b200.reconstruct import _compatible_slot_table
    from itertools import product
    N = 3
    keys = orc.basis_strings(N)
    rng = np.random.default_rng(0)
    for trial in range(3):
        ks = list(keys) if trial == 0 else [keys[i] for i in rng.permutation(len(keys))[: 27 - 5 * trial]]
        sel = _compatible_slot_table(ks, N)
        data = {k: np.zeros((1, N), np.int64) for k in ks}
        for idx, p in enumerate(product("IXYZ", repeat=N)):
            want = -2 if idx == 0 else next((i for i, k in enumerate(ks) if all(a == "I" or a == b for a, b in zip(p, k))), -1)
            assert sel[idx] == want
        if trial == 0:   # canonical order: slot = P with I->X
            for idx, p in enumerate(product("IXYZ", repeat=N)):
                if idx:
                    assert ks[sel[idx]] == "".join("X" if c == "I" else c for c in p)


def test_shard_plan_covers_everything_once():
    from ddqst_b200.distributed import plan, shard_range
    for n, w in ((6561, 8), (27, 4), (10, 3), (5, 8)):
        got = [shard_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n and all(got[i][1] == got[i + 1][0] for i in range(w - 1))
    for nb, shots, w in ((6561, 100, 8), (3, 1000, 8), (1, 999, 4), (9, 10, 2), (3, 7, 4)):
        seen = np.zeros((nb, shots), int)
        for r in range(w):
            b0, b1, s0, s1 = plan(nb, shots, r, w)
            seen[b0:b1, s0:s1] += 1
        assert (seen == 1).all(), (nb, shots, w)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from ddqst_b200.distributed import all_reduce_histograms, plan
    nb, shots, N = 5, 64, 3
    rng = np.random.default_rng(7)                                  # same table on every rank
    outcomes = rng.integers(0, 1 << N, size=(nb, shots))
    hist = torch.zeros(nb, 1 << N, dtype=torch.int32)
    b0, b1, s0, s1 = plan(nb, shots, rank, world)
    for b in range(b0, b1):
        hist[b] += torch.from_numpy(np.bincount(outcomes[b, s0:s1], minlength=1 << N)).int()
    out = all_reduce_histograms(hist.view(torch.uint32))
    flat = torch.ones(10) * (rank + 1)
    dist.all_reduce(flat)                                           # the gradient exchange of train_step
    q.put((rank, out.view(torch.int32).numpy().copy(), flat.numpy().copy()))
    dist.destroy_process_group()


def test_histogram_allreduce_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    rng = np.random.default_rng(7)
    outcomes = rng.integers(0, 8, size=(5, 64))
    want = np.stack([np.bincount(o, minlength=8) for o in outcomes])
    for _, h, flat in res:
        assert np.array_equal(h, want)                              # identical, exact counts on every rank
        assert np.allclose(flat, 3.0)
