"""Shared helpers for the parity tests: load cases and comparison metrics."""
import numpy as np

from warp3d_b200 import Material, build_structured_hex8, new_state, smooth_displacement

RTOL = 1.0e-10   # north_star: 1e-10 relative (FP64) for stresses, history, forces, stiffnesses


def relerr(a, b):
    """max |a-b| / max |b| (norm-wise relative error; 0 when both are identically zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max() if b.size else 0.0
    if scale == 0.0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / scale)


def history_views(model, hist):
    """Split a concatenated mm01 history array into (float part, integer state part)."""
    h = hist.reshape(-1, 11)
    state = h[:, 3].copy().view(np.int64) & 0xFFFFFFFF
    fl = np.delete(h, 3, axis=1)
    return fl, state


def load_path(model, nsteps=3, strain=2.0e-3, seed=1):
    """A short load history: list of (step, [(iter, u, du), ...]) with growing stretch so that a
    good fraction of the Gauss points yields from step 2 on, plus unloading in the last step."""
    rng = np.random.default_rng(seed)
    u = np.zeros(model.nodof)
    path = []
    for step in range(1, nsteps + 1):
        sgn = -0.6 if (step == nsteps and nsteps > 2) else 1.0
        target = smooth_displacement(model, strain=sgn * strain, seed=seed + step, axis=(step - 1) % 3)
        iters = []
        # iteration 0 uses du = 0-ish estimate, then two Newton-like corrections
        du0 = 0.25 * target
        du1 = 0.9 * target + rng.uniform(-1, 1, model.nodof) * 1e-6 * strain
        du2 = target
        iters.append((0, u.copy(), du0))
        iters.append((1, u.copy(), du1))
        iters.append((2, u.copy(), du2))
        path.append((step, iters))
        u = u + du2
    return path
