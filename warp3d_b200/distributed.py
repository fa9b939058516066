"""Multi-GPU layer: one process per GPU, element blocks partitioned by domain.

The reference's MPI build gives every rank the blocks whose ``elblks(2,blk) == myid``
(src/drive_eps_sig_internal_forces.f:166) and sums the internal-force vector with
``wmpi_reduce_vec(ifv, nodof)`` plus three scalar reductions and a logical OR for the step-cut flag
(:248-250, :338-340).  Here the ranks are GPUs of one box: the only exchange step of the path is
that sum, done with NCCL (``torch.distributed``) directly on the engine's device vector, on the
engine's own stream, so it is ordered after the stress kernel without a host round trip.
State never migrates; the stiffness pass needs no collective.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .model import Model, slab_domains


class _DevVec:
    """Expose an engine-owned device vector to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8", data=(int(ptr), False),
                                             version=2)


class DomainEngine:
    """Engine of one rank plus the interface-force sum across ranks."""

    def __init__(self, model: Model, rank: int, world: int, device: int = -1, group=None,
                 assign_slabs: bool = True):
        import torch
        self.torch = torch
        self.rank, self.world, self.group = rank, world, group
        if world > 1 and assign_slabs:
            model.elblks[:, 2] = slab_domains(model, world)
        self.model = model
        self.engine = Engine(model, device=device, myid=rank, numprocs=world)
        pu, pdu, pf = self.engine.device_vectors()
        dev = torch.device("cuda", torch.cuda.current_device() if device < 0 else device)
        self.stream = torch.cuda.ExternalStream(self.engine.stream, device=dev)
        self.d_ifv = torch.as_tensor(_DevVec(pf, model.nodof), device=dev)
        self.d_u = torch.as_tensor(_DevVec(pu, model.nodof), device=dev)
        self.d_du = torch.as_tensor(_DevVec(pdu, model.nodof), device=dev)
        self._scal = torch.zeros(4, dtype=torch.float64, device=dev)
        # interface dofs = dofs touched by the blocks of more than one rank (identical, sorted list
        # on every rank since the block tables are replicated): the only entries of ifv that need
        # the cross-rank sum when the caller stays on the device
        self.interface_idx = None
        if world > 1:
            off = model.block_offsets(24)
            ed = model.edest.astype(np.int64) - 1
            touched = np.zeros((world, model.nodof), dtype=bool)
            for blk in range(model.nelblk):
                touched[int(model.elblks[blk, 2]), ed[off[blk]:off[blk + 1]]] = True
            shared = np.nonzero(touched.sum(axis=0) > 1)[0]
            self.interface_idx = torch.as_tensor(shared, dtype=torch.int64, device=dev)
            self._ibuf = torch.zeros(len(shared), dtype=torch.float64, device=dev)

    # ---- device-resident Newton-iteration pieces
    def eps_sig_ifv_resident(self, step, itr, exchange="full"):
        """stress/ifv kernel for the owned blocks, then the NCCL sum of ifv over ranks.
        exchange="full": all-reduce of the whole vector (the reference's wmpi_reduce_vec; every
        rank ends with the complete ifv).  exchange="interface": only the dofs shared between
        ranks are summed (packed gather -> all-reduce -> scatter); each rank then holds the
        correct ifv on every dof its own blocks touch, which is all a device-resident caller
        needs, and the message shrinks from nodof to the slab faces."""
        self.engine.eps_sig_ifv_resident(step, itr)
        if self.world > 1:
            import torch.distributed as dist
            with self.torch.cuda.stream(self.stream):
                if exchange == "interface":
                    self.torch.index_select(self.d_ifv, 0, self.interface_idx, out=self._ibuf)
                    dist.all_reduce(self._ibuf, op=dist.ReduceOp.SUM, group=self.group)
                    self.d_ifv.index_copy_(0, self.interface_idx, self._ibuf)
                else:
                    dist.all_reduce(self.d_ifv, op=dist.ReduceOp.SUM, group=self.group)

    def tanstf_resident(self, step, itr):
        self.engine.tanstf_resident(step, itr)

    def scalars(self):
        """block scalars summed over ranks, cut flag OR-ed (wmpi_redlog / wmpi_reduce_vec)."""
        sc = self.engine.last_scalars()
        if self.world > 1:
            import torch.distributed as dist
            t = self._scal
            t.copy_(self.torch.tensor([sc["internal_energy"], sc["plastic_work"], sc["sum_ifv"],
                                       float(sc["num_term_ifv"])], dtype=self.torch.float64))
            dist.all_reduce(t, group=self.group)
            cut = self.torch.tensor([1.0 if sc["material_cut_step"] else 0.0], device=t.device)
            dist.all_reduce(cut, op=dist.ReduceOp.MAX, group=self.group)
            v = t.cpu().numpy()
            sc = dict(internal_energy=float(v[0]), plastic_work=float(v[1]), sum_ifv=float(v[2]),
                      num_term_ifv=int(round(v[3])), material_cut_step=bool(cut.item() > 0))
            if sc["material_cut_step"]:
                sc.update(internal_energy=0.0, plastic_work=0.0, sum_ifv=0.0, num_term_ifv=0)
        return sc

    # ---- the reference-facing calls with host buffers
    def drive_eps_sig_internal_forces(self, step, itr, u, du, ifv=None):
        self.engine.upload_displacements(u, du)
        self.eps_sig_ifv_resident(step, itr)
        sc = self.scalars()
        ifv = self.engine.download_ifv(ifv)
        if sc["material_cut_step"]:
            ifv[:] = 0.0
        return ifv, sc

    def tanstf(self, step, itr, out=None):
        self.engine.tanstf_resident(step, itr)
        return self.engine.download_estiff(out)

    def update(self):
        self.engine.update()

    def adaptive_save(self):
        self.engine.adaptive_save()

    def adaptive_reset(self):
        self.engine.adaptive_reset()

    def close(self):
        self.engine.close()
