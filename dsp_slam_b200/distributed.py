"""Object-sharded multi-GPU driver: one process per GPU, no data-path collective, ONE all-gather.

The reference reconstructs objects one at a time on one GPU (src/LocalMapping_util.cc:165-203).  Objects
are independent GN problems, so a batch is split into contiguous per-rank blocks (after a stable sort
by decoder class, so a rank touches as few weight sets as possible), every rank runs its block through
its own solver, and the fixed-size result records (DSPGN_RESULT_FLOATS floats per object: pose, code,
loss, status, counters) are exchanged with a single all_gather -- NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np

from . import _lib

RESULT_FLOATS = _lib.RESULT_FLOATS


def shard_plan(class_ids, world_size):
    """Returns (order, bounds): `order` = stable permutation sorting objects by class id;
    rank r owns order[bounds[r]:bounds[r+1]] (contiguous, sizes differ by at most one)."""
    class_ids = np.asarray(class_ids, dtype=np.int64)
    n = class_ids.shape[0]
    order = np.argsort(class_ids, kind="stable")
    base, extra = divmod(n, world_size)
    sizes = [base + (1 if r < extra else 0) for r in range(world_size)]
    bounds = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return order, bounds


def shard_for_rank(objs, world_size, rank):
    order, bounds = shard_plan([int(o.get("class_id", 0)) for o in objs], world_size)
    idx = order[bounds[rank]:bounds[rank + 1]]
    return [objs[i] for i in idx], idx


def all_gather_records(local, n_total, world_size, rank, order, bounds, group=None):
    """local: torch tensor (n_local, RESULT_FLOATS) on the device of the process group's backend.
    Returns a tensor (n_total, RESULT_FLOATS) in the ORIGINAL object order (on every rank)."""
    import torch
    import torch.distributed as dist
    max_n = int(np.max(np.diff(bounds)))
    pad = torch.zeros((max_n, RESULT_FLOATS), dtype=torch.float32, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world_size * max_n, RESULT_FLOATS), dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view(world_size, max_n, RESULT_FLOATS)
    res = torch.empty((n_total, RESULT_FLOATS), dtype=torch.float32, device=local.device)
    for r in range(world_size):
        k = int(bounds[r + 1] - bounds[r])
        if k:
            idx = torch.as_tensor(order[bounds[r]:bounds[r + 1]], device=local.device)
            res[idx] = out[r, :k]
    return res


def records_to_results(rec, code_len):
    """(n, RESULT_FLOATS) float32 numpy -> list of dicts like Optimizer.reconstruct_batch returns."""
    from .optimizer import ResultDict
    rec = np.ascontiguousarray(rec, dtype=np.float32)
    ints = rec.view(np.int32)
    out = []
    for i in range(rec.shape[0]):
        status = int(ints[i, 81])
        if status != 0:
            out.append(ResultDict(t_cam_obj=None, code=None, is_good=False, loss=float(rec[i, 80]), status=status))
        else:
            out.append(ResultDict(t_cam_obj=rec[i, :16].reshape(4, 4).copy(), code=rec[i, 16:16 + code_len].copy(),
                                  is_good=True, loss=float(rec[i, 80]), status=0,
                                  n_valid=int(ints[i, 82]), n_band=int(ints[i, 83])))
    return out


class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


class ShardedOptimizer:
    """reconstruct_batch over all ranks of an initialised torch.distributed process group (NCCL).
    Every rank passes the SAME full object list; every rank gets all results back."""

    def __init__(self, optimizer, group=None):
        import torch.distributed as dist
        self.opt = optimizer
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def reconstruct_batch(self, objs):
        import torch
        n = len(objs)
        order, bounds = shard_plan([int(o.get("class_id", 0)) for o in objs], self.world)
        mine = [objs[i] for i in order[bounds[self.rank]:bounds[self.rank + 1]]]
        dev = torch.device("cuda", self.opt.device)
        if mine:
            s = self.opt.solver
            s.set_stream(torch.cuda.current_stream(dev).cuda_stream)
            s.upload(mine)
            s.run(0)
            local = torch.as_tensor(_DevView(s.results_device_ptr(), len(mine) * RESULT_FLOATS), device=dev)
            local = local.view(len(mine), RESULT_FLOATS)
        else:
            local = torch.zeros((0, RESULT_FLOATS), dtype=torch.float32, device=dev)
        rec = all_gather_records(local, n, self.world, self.rank, order, bounds, self.group)
        return records_to_results(rec.cpu().numpy(), self.opt.code_len)
