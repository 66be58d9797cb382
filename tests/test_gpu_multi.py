"""Multi-GPU: a batch sharded over G devices equals the 1-GPU result bit for bit (deterministic
kernels, no reductions, no data-path collective)."""
import numpy as np
import pytest


@pytest.mark.multigpu
def test_batch_sharded_matches_single_device():
    import torch
    import phastft_b200 as pf
    G = min(torch.cuda.device_count(), 4)
    n, batch = 1 << 16, 64
    rng = np.random.default_rng(1234)
    re = rng.uniform(-1, 1, batch * n).astype(np.float32)
    im = rng.uniform(-1, 1, batch * n).astype(np.float32)
    planners = [pf.PlannerDit32(n, g) for g in range(G)]
    a_re, a_im = re.copy(), im.copy()
    pf.fft_dit_batch_sharded(a_re, a_im, pf.Direction.Forward, planners, batch)
    b_re, b_im = re.copy(), im.copy()
    pf.fft_dit_batch_sharded(b_re, b_im, pf.Direction.Forward, planners[:1], batch)
    assert np.array_equal(a_re, b_re) and np.array_equal(a_im, b_im)


@pytest.mark.gpu
def test_tables_export_import_roundtrip():
    import torch
    import phastft_b200 as pf
    n = 1 << 16
    p0 = pf.PlannerDit32(n, 0)
    p1 = pf.PlannerDit32(n, 0)
    buf = torch.empty(p0.tables_bytes(), dtype=torch.uint8, device="cuda:0")
    p0.tables_export(buf)
    buf2 = torch.zeros_like(buf)
    p1.tables_import(buf)
    p1.tables_export(buf2)
    torch.cuda.synchronize()
    assert torch.equal(buf, buf2) and int(buf.max()) > 0
    re = np.random.default_rng(0).uniform(-1, 1, n).astype(np.float32); im = re[::-1].copy()
    a, b = re.copy(), im.copy(); pf.fft_32_dit_with_planner(a, b, pf.Direction.Forward, p0)
    c, d = re.copy(), im.copy(); pf.fft_32_dit_with_planner(c, d, pf.Direction.Forward, p1)
    assert np.array_equal(a, c) and np.array_equal(b, d)
