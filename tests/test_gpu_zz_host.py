"""Host-side behaviour of the C ABI that has no counterpart in the reference's tests: page-locking helpers and
planners shared between threads (the reference's planners are plain tables behind `&`)."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_host_register_roundtrip():
    """phastft_host_register page-locks a caller-owned numpy array in place; results are unchanged and unregister succeeds."""
    import phastft_b200 as pf
    n = 1 << 16
    rng = np.random.default_rng(3)
    re = rng.uniform(-1, 1, n); im = rng.uniform(-1, 1, n)
    want = np.fft.fft(re + 1j * im)
    pf.host_register(re); pf.host_register(im)
    pf.fft_64_dit(re, im, pf.Direction.Forward)
    pf.host_unregister(re); pf.host_unregister(im)
    assert np.max(np.abs(re + 1j * im - want)) / np.max(np.abs(want)) <= 4 * 2.0 ** -52 * 16


@pytest.mark.gpu
def test_one_planner_shared_by_threads():
    """The reference's planners are shared by `&` between threads (they are plain tables); here a planner owns device
    staging buffers and streams, so every *_host call holds the plan for its whole duration.  Eight threads, one
    planner per API family, different data each: every result must be right."""
    import threading
    import phastft_b200 as pf
    n = 1 << 15
    rng = np.random.default_rng(11)
    p_c2c = pf.PlannerDit64(n); p_r2c = pf.PlannerR2c64(n)
    jobs, errs = [], []
    for t in range(8):
        re = rng.uniform(-1, 1, n); im = rng.uniform(-1, 1, n)
        jobs.append((re, im, np.fft.fft(re + 1j * im), np.fft.rfft(re)))

    def work(t):
        re, im, want, rwant = jobs[t]
        try:
            for _ in range(10):
                a, b = re.copy(), im.copy()
                pf.fft_64_dit_with_planner(a, b, pf.Direction.Forward, p_c2c)
                if np.max(np.abs(a + 1j * b - want)) / np.max(np.abs(want)) > 1e-13:
                    errs.append(("c2c", t))
                ore = np.zeros(n // 2 + 1); oim = np.zeros(n // 2 + 1)
                pf.r2c_fft_f64_with_planner(re, ore, oim, p_r2c)
                if np.max(np.abs(ore + 1j * oim - rwant)) / np.max(np.abs(rwant)) > 1e-11:
                    errs.append(("r2c", t))
                back = np.zeros(n)
                pf.c2r_fft_f64_with_planner(ore, oim, back, p_r2c)
                if np.max(np.abs(back - re)) > 1e-11:
                    errs.append(("c2r", t))
                z = (re + 1j * im).astype(np.complex128)
                pf.fft_64_interleaved_with_planner(z, pf.Direction.Forward, p_c2c)
                if np.max(np.abs(z - want)) / np.max(np.abs(want)) > 1e-13:
                    errs.append(("interleaved", t))
        except Exception as e:  # noqa: BLE001
            errs.append((repr(e), t))

    threads = [threading.Thread(target=work, args=(t,), daemon=True) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(120)
    assert not any(th.is_alive() for th in threads), "a host call never returned (lock-up inside the library)"
    assert not errs, errs[:5]
