"""Host-side helpers for distributed runs (one process per GPU, ``torch.distributed`` as plumbing).

* ``peer_segments``: the per-neighbour-rank segments of the MPI interface list (the analogue of
  ``mpi_neighbor_ranks`` / ``mpi_neighbor_interfaces`` in ``P4estMPICache`` dgsem_p4est/dg_parallel.jl:8-20,
  built by ``init_mpi_neighbor_connectivity`` :185-236).
* ``HostHaloExchange``: face-state exchange on host arrays with isend/irecv -- what the reference does
  with MPI (``start_mpi_send!``/``finish_mpi_receive!`` dg_parallel.jl:66-182).  It is used with the gloo
  backend to test the partition/connectivity logic on CPU; on GPUs the exchange runs inside
  libtrixi_b200 (pack kernels store directly into the peer's receive buffer over NVLink).
* ``allreduce_min``: the ``MPI.Allreduce!(dt, min)`` of the distributed ``max_dt`` (stepsize_dg3d.jl:264-279).
"""
from __future__ import annotations

import numpy as np


def peer_segments(neighbor_ranks):
    """[(peer, offset, count)] for an MPI interface list sorted by neighbour rank."""
    neighbor_ranks = np.asarray(neighbor_ranks)
    if neighbor_ranks.size == 0:
        return []
    peers, offsets, counts = np.unique(neighbor_ranks, return_index=True, return_counts=True)
    return [(int(p), int(o), int(c)) for p, o, c in zip(peers, offsets, counts)]


class HostHaloExchange:
    def __init__(self, semi, dist):
        import torch
        self.torch, self.dist = torch, dist
        mi = semi.cache.mpi_interfaces
        self.segments = peer_segments(mi.neighbor_ranks)
        self.local_side = mi.local_sides - 1  # 0-based slot of the local data in mpi_u[2, ...]
        self.nv = semi.equations.nvars
        self.nf = semi.solver.nnodes ** (semi.mesh.ndims - 1)

    def exchange(self, mpi_u_flat):
        """Fill the remote side of mpi_u[2, nv, nf, MI] (Fortran order, flat) from the neighbours."""
        torch, dist = self.torch, self.dist
        nmi = self.local_side.shape[0]
        if nmi == 0:
            return
        mu = mpi_u_flat.reshape((2, self.nv, self.nf, nmi), order="F")
        idx = np.arange(nmi)
        local = mu[self.local_side, :, :, idx]  # [MI, nv, nf]
        reqs, recvs = [], []
        for peer, off, cnt in self.segments:
            send = torch.from_numpy(np.ascontiguousarray(local[off:off + cnt]))
            recv = torch.empty_like(send)
            reqs.append(dist.isend(send, dst=peer))
            reqs.append(dist.irecv(recv, src=peer))
            recvs.append((off, cnt, recv, send))
        for r in reqs:
            r.wait()
        for off, cnt, recv, _ in recvs:
            mu[1 - self.local_side[off:off + cnt], :, :, idx[off:off + cnt]] = recv.numpy()


def allreduce_min(value, dist):
    """min over ranks that propagates NaN like the reference's `min` (a NaN dt must abort every rank, not just the
    one that produced it: the others would otherwise wait in the halo exchange forever)."""
    import math

    import torch
    v = float(value)
    t = torch.tensor([-math.inf if math.isnan(v) else v], dtype=torch.float64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out = float(t.item())
    return math.nan if out == -math.inf else out
