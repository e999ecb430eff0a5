"""Multi-GPU plumbing for the hot path: one process per GPU, tiles sharded by rank, and exactly ONE collective --
a broadcast of the checkpoint's flat fp32 state (157.8 MB) from rank 0 over NCCL / NVLink.

Replaces the reference's per-forward `nn.DataParallel` replicate / scatter / gather
(robosat/tools/predict.py:63): every rank owns its tiles end to end.
"""

from collections import OrderedDict

import torch
import torch.distributed as dist


def unet_state_template(num_classes, prefix="module."):
    """names -> (shape, dtype) of the reference checkpoint's state_dict, known on every rank without the file"""
    from robosat_b200 import synth

    t = OrderedDict()
    for name, shape, kind in synth.unet_param_shapes(num_classes):
        t[prefix + name] = (tuple(shape), torch.int64 if kind == "bn_count" else torch.float32)
    return t


def broadcast_state_dict(sd, template, device, src=0):
    """Rank `src` passes its state_dict, the others pass None; everyone returns the same dict (tensors on `device`).

    One flat fp32 buffer + one small int64 buffer -> two `dist.broadcast` calls issued back to back (the int64 one is
    53 scalars); `template` maps names to (shape, dtype) or to example tensors."""
    spec = OrderedDict()
    for k, v in template.items():
        spec[k] = (tuple(v.shape), v.dtype) if torch.is_tensor(v) else v
    n_f = sum(int(torch.Size(s).numel()) for s, d in spec.values() if d == torch.float32)
    n_i = sum(int(torch.Size(s).numel()) for s, d in spec.values() if d == torch.int64)
    flat_f = torch.empty(n_f, dtype=torch.float32, device=device)
    flat_i = torch.empty(max(n_i, 1), dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        assert sd is not None and list(sd.keys()) == list(spec.keys()), "state_dict does not match the template"
        of = oi = 0
        for k, (shape, dtype) in spec.items():
            n = int(torch.Size(shape).numel())
            if dtype == torch.float32:
                flat_f[of:of + n].copy_(sd[k].reshape(-1))
                of += n
            else:
                flat_i[oi:oi + n].copy_(sd[k].reshape(-1))
                oi += n
    dist.broadcast(flat_f, src=src)
    dist.broadcast(flat_i, src=src)
    out = OrderedDict()
    of = oi = 0
    for k, (shape, dtype) in spec.items():
        n = int(torch.Size(shape).numel())
        if dtype == torch.float32:
            out[k] = flat_f[of:of + n].view(shape)
            of += n
        else:
            out[k] = flat_i[oi:oi + n].view(shape)
            oi += n
    return out


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for this rank (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(flat, world=None):
    """In-place sum of one flat tensor over all ranks: the single data-path collective of data-parallel training
    (replaces DataParallel's per-step reduce-add of gradients to GPU 0). The caller scales the loss by
    1 / world beforehand, so the summed gradient is the global mean and no extra elementwise pass is needed."""
    if dist.is_available() and dist.is_initialized() and (world or dist.get_world_size()) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def check_async_error(device, what):
    """NCCL collectives are asynchronous: a failure (peer died, link error) surfaces only when the stream is synchronised.
    Called after each collective phase of the tools so that a broken rank turns into a non-zero exit at a named place
    instead of a hang or a silently wrong result. TORCH_NCCL_ASYNC_ERROR_HANDLING=1 (set by the tools) additionally lets
    the NCCL watchdog abort the process on a timed-out collective."""
    try:
        torch.cuda.synchronize(device)
    except RuntimeError as exc:  # pragma: no cover  (needs a real fault)
        raise SystemExit("Error: %s failed on rank %d: %s" % (what, dist.get_rank() if dist.is_initialized() else 0, exc))
