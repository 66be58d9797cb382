"""Multi-GPU: a batch sharded over G devices equals the 1-GPU result bit for bit (deterministic
kernels, no reductions, no data-path collective)."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_batch_sharded_matches_single_device():
    """Runs with however many devices the box has (the driver's `-m gpu` box has one: the sharded entry point then degenerates
    to one shard, which still exercises the partition and the host pipeline; with 2-8 devices it is the cross-device check)."""
    import torch
    import phastft_b200 as pf
    G = max(1, min(torch.cuda.device_count(), 8))
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


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [False, True])
def test_batch_host_pipeline_chunks_and_stride(pinned, monkeypatch):
    """The host-resident batch goes through a 3-slot H2D / FFT / D2H pipeline: many small chunks, a ragged last
    chunk, a batch stride larger than N (gaps must stay untouched), pageable and page-locked memory."""
    import torch
    import phastft_b200 as pf
    n, batch, stride = 1 << 14, 37, (1 << 14) + 24
    monkeypatch.setenv("PHASTFT_HOST_CHUNK_MB", "1")          # 1 MiB / (2^14 * 8 B) = 8 transforms per chunk -> 5 chunks
    rng = np.random.default_rng(99)
    total = (batch - 1) * stride + n
    t_re = torch.from_numpy(rng.uniform(-1, 1, total)); t_im = torch.from_numpy(rng.uniform(-1, 1, total))
    if pinned:
        t_re, t_im = t_re.pin_memory(), t_im.pin_memory()
    re, im = t_re.numpy(), t_im.numpy()
    ref_re, ref_im = re.copy(), im.copy()
    planner = pf.PlannerDit64(n, 0)
    pf.fft_dit_batch_sharded(re, im, pf.Direction.Forward, [planner], batch, batch_stride=stride)
    tol = 4 * 2.0 ** -52 * 14
    for b in range(batch):
        want = np.fft.fft(ref_re[b * stride:b * stride + n] + 1j * ref_im[b * stride:b * stride + n])
        got = re[b * stride:b * stride + n] + 1j * im[b * stride:b * stride + n]
        assert np.max(np.abs(got - want)) / np.max(np.abs(want)) <= tol, b
        if b + 1 < batch:   # the gap between signals is not part of any transform
            assert np.array_equal(re[b * stride + n:(b + 1) * stride], ref_re[b * stride + n:(b + 1) * stride])
            assert np.array_equal(im[b * stride + n:(b + 1) * stride], ref_im[b * stride + n:(b + 1) * stride])
    # and back: forward then reverse is the identity
    pf.fft_dit_batch_sharded(re, im, pf.Direction.Reverse, [planner], batch, batch_stride=stride)
    assert np.max(np.abs(re - ref_re)) <= 64 * tol and np.max(np.abs(im - ref_im)) <= 64 * tol
