"""One-process-per-GPU cooperation (torch.distributed): host-side merge logic of the sharded enumeration.

Every rank calls the device enumerator with shard=(rank, world) on the SAME inputs (fplll_b200.enumeration.enumerate_svp),
then `merge_enum_results` makes all ranks agree on one result: node counts are summed, the shortest vector wins with a
deterministic tie-break (smaller dist, then lexicographically smaller coefficient vector) so the replicated BKZ control
flow stays identical on every rank (SURVEY §8e "deterministic mode").  The GSO path itself never communicates
(replicas only).  Backend: NCCL on GPUs, gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def merge_enum_results(local, d, group=None, device=None):
    """local: dict(solutions=[(dist, x)...], nodes=uint64[d], stats={...}) from enumerate_svp on this rank's shard."""
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device("cpu")
    nodes = torch.as_tensor(np.asarray(local["nodes"], dtype=np.int64), device=dev)
    dist.all_reduce(nodes, op=dist.ReduceOp.SUM, group=group)
    best = torch.full((d + 1,), float("inf"), dtype=torch.float64, device=dev)
    if local["solutions"]:
        dd, x = local["solutions"][-1]
        best[0] = dd
        best[1:] = torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
    gathered = [torch.empty_like(best) for _ in range(world)]
    dist.all_gather(gathered, best, group=group)
    cands = [g.cpu().numpy() for g in gathered if np.isfinite(g[0].item())]
    sols = []
    if cands:
        cands.sort(key=lambda c: (c[0], tuple(c[1:])))
        sols = [(float(cands[0][0]), cands[0][1:].copy())]
    leaves = torch.tensor([int(local["stats"].get("leaves", 0))], dtype=torch.int64, device=dev)
    dist.all_reduce(leaves, op=dist.ReduceOp.SUM, group=group)
    return dict(solutions=sols, nodes=nodes.cpu().numpy().astype(np.uint64), leaves=int(leaves.item()))


_attached = {}


def attach_peers(group=None, device_index=0):
    """Once per job: let the ranks' enumerators reach each other's radius words over NVLink (CUDA IPC; include/b200enum.h
    b200enum_ipc_*).  The 64-byte handles are all-gathered through the process group (NCCL on GPUs).  Without it sharded
    calls run with private radii."""
    from . import enumeration as en
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    key = (id(group), device_index)
    if world < 2 or _attached.get(key):
        return bool(_attached.get(key))
    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device_index) if nccl else torch.device("cpu")
    mine = torch.frombuffer(bytearray(en.ipc_export(device_index)), dtype=torch.uint8).to(dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    en.ipc_attach(device_index, world, rank, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    dist.barrier(group=group)
    _attached[key] = True
    return True


def enumerate_svp_distributed(mut, rdiag, pruning, maxdist, group=None, device_index=0, fixed_radius=False):
    """sharded enumeration across the ranks of `group` (one GPU per rank): the subtree roots are dealt round-robin in
    order of promise; after attach_peers() the ranks push radius improvements to each other inside the kernel (NVLink
    peer memory).  The result merge below (NCCL) is also what keeps call k+1 from starting before every rank finished
    call k."""
    from . import enumeration as en
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = en.enumerate_svp(mut, rdiag, pruning, maxdist, fixed_radius=fixed_radius, devices=[device_index],
                             shard=(rank, world))
    dev = torch.device("cuda", device_index) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    return merge_enum_results(local, len(rdiag), group=group, device=dev)
