"""Multi-GPU plumbing: one process per GPU, events and injections sharded contiguously (optionally hyper-points
too: `grid_coords`), one all-reduce of the per-rank partials.

The split rule is the reference's (CHIMERA/parallel.py:68-73, 94-99: `n // R` per rank, the first
`n % R` ranks get one more).  The exchanged payload is `(n_hyper, 3)` f64 =
[sum_local log L_i, sum_local w_inj, sum_local w_inj^2]; the reference all-reduces a zero-padded
`(nparams, tot_inj)` matrix instead (parallel.py:289-294)."""
import numpy as np


def shard_bounds(n, rank, world):
  """[lo, hi) of `rank`'s contiguous chunk of `n` items."""
  base, rem = divmod(int(n), int(world))
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def grid_coords(rank, world, hyper_groups=1):
  """2-D layout for large walker batches (the reference's 'both' scheme, CHIMERA/parallel.py:132-229): the world is
  `hyper_groups` groups of `world // hyper_groups` ranks.  Returns (event_shard, n_event_shards, hyper_group).
  A rank owns event/injection shard `event_shard` and evaluates only the hyper-points of `hyper_group`
  (`shard_bounds(n_hyper, hyper_group, hyper_groups)`); every (hyper-point, event) unit and every
  (hyper-point, injection) pair is evaluated by exactly one rank, so ONE SUM all-reduce of the zero-filled
  (n_hyper, 3) partials over the whole world still gives the global sums."""
  k = int(hyper_groups)
  if k < 1 or world % k != 0:
    raise ValueError(f"hyper_groups={k} must divide the world size {world}")
  E = world // k
  return rank % E, E, rank // E


def dist_info(group=None):
  """(rank, world) of the torch.distributed group, or (0, 1) when not initialised."""
  try:
    import torch.distributed as dist
  except Exception:
    return 0, 1
  if not (dist.is_available() and dist.is_initialized()):
    return 0, 1
  return dist.get_rank(group), dist.get_world_size(group)


def backend(group=None):
  """Name of the torch.distributed backend in use ('' when not initialised)."""
  try:
    import torch.distributed as dist
  except Exception:
    return ""
  if not (dist.is_available() and dist.is_initialized()):
    return ""
  return dist.get_backend(group)


def allgather_counts(n_local, group=None):
  """Per-rank item counts (list of ints) for pre-sharded inputs."""
  import torch
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return [int(n_local)]
  t = torch.tensor([int(n_local)], dtype=torch.int64)
  if dist.get_backend(group) == "nccl":
    t = t.cuda()
  outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
  dist.all_gather(outs, t, group=group)
  return [int(o.item()) for o in outs]


def allreduce_partials(partials, group=None):
  """SUM all-reduce of the partials over ranks.  Accepts a NumPy array (moved through a CPU
  tensor: gloo, or NCCL via a staging copy to the current CUDA device) or a torch tensor that
  already lives on the right device (NCCL over NVLink, no host round trip).  In-place for tensors."""
  import torch
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return partials
  if isinstance(partials, torch.Tensor):
    dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials
  t = torch.from_numpy(np.ascontiguousarray(partials, dtype=np.float64))
  if dist.get_backend(group) == "nccl":
    t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
  dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
  return t.numpy()


def allgather_events(local, counts, group=None):
  """Concatenate per-rank `(n_hyper, Nev_local)` blocks along the event axis (only used by
  `compute_all`, which returns per-event values)."""
  import torch
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return local
  world = dist.get_world_size(group)
  nh = local.shape[0]
  mx = max(counts)
  pad = np.zeros((nh, mx))
  pad[:, :local.shape[1]] = local
  t = torch.from_numpy(pad)
  nccl = dist.get_backend(group) == "nccl"
  if nccl:
    t = t.cuda()
  outs = [torch.empty_like(t) for _ in range(world)]
  dist.all_gather(outs, t, group=group)
  return np.concatenate([o.cpu().numpy()[:, :c] for o, c in zip(outs, counts)], axis=1)
