"""Joining the solvers of several ranks (one GPU each) into one image-sharded problem.

``torch.distributed`` is plumbing only: it carries the 128-byte NCCL id, the 64-byte CUDA IPC handles of
the peer-memory blocks and one element count.  The communicator and the mapped peer blocks are created
once per (process group, device) and shared by every solver that attaches afterwards."""

import os

from . import _lib

_COMMS = {}     # (process group, device, world size) -> _lib.Comm, created once per process


def attach(handle, device, nx_local, dist, group=None):
    """Attach `handle` to the group's communicator.  Returns (world size, peer-memory exchange in use).

    The per-iteration sums are then reduced over the ranks on the device: inside the scalar kernel over
    peer memory where the ranks can map each other's memory (one node, <= 8 ranks), by a small NCCL
    all-reduce otherwise.  Every rank must take the same decision, hence the all-reduce of the outcome."""
    import torch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device('cuda', device)
    key = (id(group), device, world)
    comm = _COMMS.get(key)
    if comm is None:
        nccl_lib = _lib.nccl_library_path()
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(_lib.comm_unique_id(nccl_lib)), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, src=0, group=group)
        comm = _lib.Comm(nccl_lib, bytes(uid.cpu().numpy().tobytes()), rank, world, device)
        _COMMS[key] = comm
    nx = torch.tensor([float(nx_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(nx, group=group)
    handle.attach_comm(comm, float(nx.item()))
    p2p = False
    if 1 < world <= 8 and os.environ.get('SPCSC_P2P', '1') != '0':
        if comm.p2p is not None:
            # an earlier solver of this group has settled it (the blocks live with the communicator):
            # no handle exchange, no collective
            if comm.p2p:
                p2p = bool(handle.p2p_attach(rank, world, None))
            return world, p2p
        # every rank takes part in every collective below, whatever happened locally: a rank whose
        # export failed sends a zero handle and votes "no"
        ok = True
        try:
            raw = bytearray(handle.p2p_export())
        except _lib.SpcscError:
            raw, ok = bytearray(64), False
        mine = torch.frombuffer(raw, dtype=torch.uint8).to(dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if flag.item() == 1.0:
            blob = b''.join(bytes(t.cpu().numpy().tobytes()) for t in allh)
            try:
                ok = bool(handle.p2p_attach(rank, world, blob))
            except _lib.SpcscError:
                ok = False
        else:
            ok = False
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if flag.item() != 1.0 and ok:
            handle.p2p_detach()
        p2p = bool(flag.item() == 1.0)
        comm.p2p = p2p
    return world, p2p
