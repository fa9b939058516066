"""Multi-GPU layer: one process per GPU, element blocks partitioned by domain.

The reference's MPI build gives every rank the blocks whose ``elblks(2,blk) == myid``
(src/drive_eps_sig_internal_forces.f:166) and sums the internal-force vector with
``wmpi_reduce_vec(ifv, nodof)`` plus three scalar reductions and a logical OR for the step-cut flag
(:248-250, :338-340).  Here the ranks are GPUs of one box and the exchange step lives BEHIND THE
C ABI (include/w3d_engine.h: w3d_comm_unique_id / w3d_comm_init): NCCL send/receive pairs between
neighbouring domains (or one all-reduce of the whole vector, the reference's semantics) and one
5-double all-reduce for the scalars, all on the engine's own stream.  A Fortran MPI caller reaches
exactly the same code.  This module only does what the MPI layer of the caller would do: it hands
rank 0's NCCL id to the other ranks (``torch.distributed`` object broadcast; any transport would
do) and assigns contiguous block slabs to ranks.  State never migrates; the stiffness pass needs
no collective.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine, comm_owned, comm_unique_id
from .model import Model, slab_domains


def share_unique_id(rank: int, world: int, group=None):
    """rank 0 creates the NCCL id through the C ABI, everybody receives it."""
    if world == 1:
        return comm_unique_id()
    import torch.distributed as dist
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


class DomainEngine:
    """Engine of one rank; the cross-rank sum of the internal forces happens inside the engine."""

    def __init__(self, model: Model, rank: int, world: int, device: int = -1, group=None,
                 assign_slabs: bool = True, exchange: str = "neighbours", bcast_inputs: bool = False):
        self.rank, self.world, self.group = rank, world, group
        if world > 1 and assign_slabs:
            model.elblks[:, 2] = slab_domains(model, world)
        self.model = model
        self.engine = Engine(model, device=device, myid=rank, numprocs=world)
        self.bcast_inputs = bool(bcast_inputs) and world > 1
        if world > 1:
            self.engine.comm_init(share_unique_id(rank, world, group), rank, world)
            self.engine.comm_set_options(exchange, self.bcast_inputs)
        self._owned = None

    # ---- which part of the full vector this rank returns
    def owned_mask(self):
        """dofs this rank owns (lowest rank whose blocks touch a dof): the full ifv is the
        concatenation of every rank's owned entries."""
        if self._owned is None:
            self._owned = (comm_owned(self.model, self.rank, self.world) if self.world > 1
                           else np.ones(self.model.nodof, dtype=bool))
        return self._owned

    def owned_range(self):
        """(first 1-based, n) of the smallest contiguous dof range holding every owned dof; for slab
        partitions of a structured mesh the range holds nothing else."""
        idx = np.nonzero(self.owned_mask())[0]
        if len(idx) == 0:
            return 1, 0
        return int(idx[0]) + 1, int(idx[-1] - idx[0]) + 1

    # ---- device-resident Newton-iteration pieces
    def eps_sig_ifv_resident(self, step, itr, exchange=None):
        """stress / ifv kernels of the owned blocks, then the exchange step, inside the engine."""
        self.engine.eps_sig_ifv_resident(step, itr)

    def tanstf_resident(self, step, itr):
        self.engine.tanstf_resident(step, itr)

    def scalars(self):
        """block scalars summed over ranks, cut flag OR-ed: already done by the engine."""
        return self.engine.last_scalars()

    # ---- the reference-facing calls with host buffers
    def drive_eps_sig_internal_forces(self, step, itr, u, du, ifv=None, owned_only=False):
        """ifv: the full vector on this rank (every rank holds every dof its blocks touch; with
        exchange="full" all of it).  owned_only: copy back only this rank's owned range."""
        if self.bcast_inputs and self.rank != 0:
            u = du = None
        if owned_only:
            _, sc = self.engine.drive_eps_sig_internal_forces(step, itr, u, du, ifv=False)
            first, n = self.owned_range()
            out = ifv[first - 1:first - 1 + n] if ifv is not None else None
            out = self.engine.download_ifv_range(first, n, out)
            if sc["material_cut_step"]:
                out[:] = 0.0
            return out, sc
        return self.engine.drive_eps_sig_internal_forces(step, itr, u, du, ifv)

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
