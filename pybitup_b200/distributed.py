"""One process per GPU: chain / sample sharding and the small moment all-reduce.

Chains (and propagation samples) are independent units, so each rank owns a contiguous range of the global index
space and the data path has NO collective.  The only exchange is a sum over ranks of a few hundred doubles (chain
sums, propagation histograms) through ``torch.distributed`` -- NCCL over NVLink on the GPU box, gloo in the CPU
tests.  It replaces the blocking pickled ``comm.send`` / ``comm.recv`` of whole evaluation matrices in the reference
(``pybitup/solve_problem.py:578-592``), the only message passing that code has.
"""
import numpy as np

from . import diagnostics


def local_device():
    """The CUDA device of this process: LOCAL_RANK under torchrun (one process per GPU), else torch's current device."""
    import os
    import torch
    if "LOCAL_RANK" in os.environ and torch.cuda.is_available():
        return int(os.environ["LOCAL_RANK"]) % max(torch.cuda.device_count(), 1)
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def ensure_process_group():
    """Under torchrun (WORLD_SIZE > 1 in the environment) with no process group yet: bind this process to its LOCAL_RANK
    GPU and initialise NCCL (gloo without CUDA: the CPU tests).  A no-op for a plain single-process run."""
    import os
    import torch
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or not dist.is_available() or dist.is_initialized():
        return
    if torch.cuda.is_available():
        dev = local_device()
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")


def shard_range(n_total, world_size, rank):
    """Contiguous [first, first + count) of n_total units for `rank`; the remainder goes to the low ranks."""
    base, rem = divmod(int(n_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class Reducer:
    """reduce(array, op) over the ranks of a torch.distributed group (op: 'sum' | 'min' | 'max')."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        if device is None and self.active and dist.get_backend(group) == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())
        elif isinstance(device, int):
            device = torch.device("cuda", device)
        self.device = device if device is not None else torch.device("cpu")

    def __call__(self, a, op="sum"):
        a = np.ascontiguousarray(a)
        if not self.active or self.world_size == 1:
            return a
        dist = self._dist
        t = self._torch.from_numpy(a.copy()).to(self.device)
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op],
                        group=self.group)
        return t.cpu().numpy().astype(a.dtype, copy=False).reshape(a.shape)

    def broadcast0(self, a):
        """The array of rank 0 on every rank (small host arrays: the final state of global chain 0, file decisions)."""
        a = np.ascontiguousarray(a)
        if not self.active or self.world_size == 1:
            return a
        t = self._torch.from_numpy(a.copy()).to(self.device)
        self._dist.broadcast(t, src=self._dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        return t.cpu().numpy().astype(a.dtype, copy=False).reshape(a.shape)

    @property
    def device_sum(self):
        """In-place sum of a CUDA tensor over the ranks on the current stream (NCCL), or None when the group cannot do
        that (gloo, a single rank): callers then go through __call__ with host arrays."""
        if not self.active or self.world_size == 1 or self._dist.get_backend(self.group) != "nccl":
            return None

        def f(t):
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
            return t
        return f


def global_chain_statistics(local_sums_fn, d, reducer=None, mu0=None):
    """Pooled mean / covariance / R-hat over the chains of all ranks.

    local_sums_fn(mu) -> this rank's chain sums about mu (``ChainEngine.chain_sums`` on the GPU).  Two small
    all-reduces: the first locates the pooled mean, the second takes the second moments about it, which removes the
    cancellation a one-pass sum of squares would have."""
    reducer = reducer or (lambda a, op="sum": a)
    mu = np.zeros(d) if mu0 is None else np.asarray(mu0, dtype=np.float64)
    s = reducer(local_sums_fn(mu), "sum")
    lay, _ = diagnostics.sums_layout(d)
    mu = mu + s[lay["n_delta"]] / s[1]
    if hasattr(reducer, "active") and reducer.active:
        # every rank must use the bit-identical reference point
        mu = reducer(mu, "max")
    s = reducer(local_sums_fn(mu), "sum")
    return diagnostics.finalize(s, mu, d)
