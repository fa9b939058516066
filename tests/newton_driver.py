"""A minimal global Newton driver (test infrastructure) around the block hot path, modelled on
mnralg (src/mnralg.f:20-620): iteration 0 stress pass + tangent, then solve / stress update /
residual / convergence test per iteration, n+1 -> n rollover on convergence.  The backend is either
the CPU oracle or the CUDA engine; both see identical inputs, so iteration counts and residual
norms can be compared (north_star: identical counts, norms to 1e-8)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def icp_table(n=24):
    """(row, col) 0-based for each of the n(n+1)/2 upper-triangle terms (setup.f:268-276)."""
    rows, cols = [], []
    for j in range(n):
        for i in range(j + 1):
            rows.append(i); cols.append(j)
    return np.array(rows), np.array(cols)


def assemble(model, estiff):
    """EBE upper triangles (300, span) -> global symmetric CSR via edest."""
    ek = estiff.reshape(-1, 300)
    ed = model.edest.reshape(-1, 24).astype(np.int64) - 1
    r, c = icp_table()
    gi = ed[:, r].ravel(); gj = ed[:, c].ravel(); v = ek.ravel()
    off = gi != gj
    I = np.concatenate([gi, gj[off]]); J = np.concatenate([gj, gi[off]]); V = np.concatenate([v, v[off]])
    return sp.csr_matrix((V, (I, J)), shape=(model.nodof, model.nodof))


class OracleBackend:
    def __init__(self, model, state):
        from oracle.oracle_py import Oracle
        self.o, self.s = Oracle(model), state

    def eps_sig_ifv(self, step, it, u, du):
        return self.o.eps_sig_ifv(self.s, step, it, u, du)

    def tanstf(self, step, it, u, du):
        return self.o.tanstf(self.s, step, it, u, du).copy()

    def update(self):
        self.o.update(self.s); self.o.adaptive_save(self.s)


class EngineBackend:
    def __init__(self, model):
        from warp3d_b200.engine import Engine
        self.e = Engine(model)

    def eps_sig_ifv(self, step, it, u, du):
        return self.e.drive_eps_sig_internal_forces(step, it, u, du)

    def tanstf(self, step, it, u, du):
        return self.e.tanstf(step, it, u, du)

    def update(self):
        self.e.update(); self.e.adaptive_save()


def solve_steps(model, backend, fixed_dofs, fixed_increments, nsteps, tol=1e-10, maxit=12):
    """Displacement-controlled load steps.  fixed_increments[step] gives the imposed increment
    on fixed_dofs.  Returns (u, log) where log[step] = list of (residual norm, reaction norm) per iteration."""
    n = model.nodof
    free = np.setdiff1d(np.arange(n), fixed_dofs)
    u = np.zeros(n)
    log = []
    for step in range(1, nsteps + 1):
        du = np.zeros(n)
        ifv, _ = backend.eps_sig_ifv(step, 0, u, du)
        K = assemble(model, backend.tanstf(step, 0, u, du))
        norms = []
        for it in range(1, maxit + 1):
            delta = np.zeros(n)
            if it == 1:
                delta[fixed_dofs] = fixed_increments[step - 1]
            rhs = (-ifv)[free] - K[free][:, fixed_dofs] @ delta[fixed_dofs]
            delta[free] = spla.spsolve(K[free][:, free].tocsc(), rhs)
            du = du + delta
            ifv, sc = backend.eps_sig_ifv(step, it, u, du)
            assert not sc["material_cut_step"]
            rn = float(np.linalg.norm(ifv[free]))
            ref = float(np.linalg.norm(ifv[fixed_dofs])) or 1.0
            norms.append((rn, ref))
            if rn <= tol * ref:
                break
            K = assemble(model, backend.tanstf(step, it, u, du))
        log.append(norms)
        u = u + du
        backend.update()
    return u, log
