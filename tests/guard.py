"""Guard-band check for sampled bits (test infrastructure).

north_star: "bit-exact sampled bitstrings ... when fed identical random streams".  Every draw of the reverse sampler is
``bit = u*(p0+p1) < p1`` where (p0, p1) is a deterministic function of the denoiser logits of that (shot, qubit) and
``u`` comes from the shared Philox stream.  The CUDA path and the CPU oracle compute the logits with different
summation orders (and, in production mode, in bf16), so a draw may legitimately differ only where ``u`` sits so close to
the decision threshold that the *measured* logit difference of that very element can move the threshold across it.

``assert_draws_in_guard_band`` makes that exact: for every element it evaluates the ORACLE's reverse step twice, on the
oracle's logits pushed by the element's own measured error (plus a small slack for the post-logit fp32 arithmetic:
exp / division order) towards bit 0 and towards bit 1.  The decision depends on the logits only through d = l1 - l0 and
is monotone in d, so the draw the CUDA path made from ITS logits must equal one of the two.  Any bit that equals neither
is outside the band: a real bug (wrong qubit lane, wrong table row, wrong Philox counter, wrong posterior arithmetic),
and the assertion fails on the first one.  The number of draws that fall inside the band is returned so the caller can
bound it (a band wide enough to swallow everything would make the test vacuous).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ddqst_oracle as orc


def oracle_reverse_step(mode, logits, x_t, t, betas, Q, seed, basis, shot_offset, num_qubits):
    """x_{t-1} the oracle draws from ``logits`` (RQC/diffusion.py:62-79 or SS/diffusion.py:67-80) for shots
    shot_offset + [0, B)."""
    B = x_t.shape[0]
    idx = shot_offset + np.arange(B)
    if mode == "posterior":
        u = torch.from_numpy(orc.stream_uniforms(seed, basis, t, orc.SITE_POSTERIOR, idx, num_qubits))
        return orc.posterior_step(logits, x_t, betas[t], Q[t - 1], u)[0]
    probs = F.softmax(logits, dim=2)
    u = torch.from_numpy(orc.stream_uniforms(seed, basis, t, orc.SITE_X0HAT, idx, num_qubits))
    x0_hat = orc.draw_bits(probs, u)
    if t > 1:
        return orc.q_sample_marginal(Q, x0_hat, torch.full((B,), t - 1, dtype=torch.long), seed, basis,
                                     site=orc.SITE_RENOISE, t_field=t, row_offset=shot_offset)
    return x0_hat


def assert_draws_in_guard_band(got_bits, got_logits, want_logits, mode, x_t, t, betas, Q, seed, basis, shot_offset,
                               num_qubits, slack=2e-6, what=""):
    """-> (n_mismatch_vs_nominal, n_inside_band, n_total).  Raises AssertionError if any draw is outside the band."""
    got_bits, got_logits, want_logits = got_bits.cpu(), got_logits.cpu().float(), want_logits.float()
    e = (got_logits - want_logits).abs() + slack * want_logits.abs().clamp(min=1.0)
    to0, to1 = want_logits.clone(), want_logits.clone()
    to0[..., 0] += e[..., 0]; to0[..., 1] -= e[..., 1]
    to1[..., 0] -= e[..., 0]; to1[..., 1] += e[..., 1]
    step = lambda lg: oracle_reverse_step(mode, lg, x_t, t, betas, Q, seed, basis, shot_offset, num_qubits)
    nominal, x_a, x_b = step(want_logits), step(to0), step(to1)
    inside = x_a != x_b
    ok = (got_bits == x_a) | (got_bits == x_b)
    if not bool(ok.all()):
        bad = torch.nonzero(~ok)
        r, q = (int(v) for v in bad[0])
        raise AssertionError(
            f"{what}: {bad.shape[0]} draw(s) OUTSIDE the guard band at t={t}; first: shot {shot_offset + r} qubit {q}: "
            f"got {int(got_bits[r, q])}, oracle {int(nominal[r, q])}, logits got {got_logits[r, q].tolist()} want {want_logits[r, q].tolist()}")
    mism = got_bits != nominal
    assert bool((inside | ~mism).all())          # every mismatch is a band draw (implied by the above; kept explicit)
    return int(mism.sum()), int(inside.sum()), got_bits.numel()
