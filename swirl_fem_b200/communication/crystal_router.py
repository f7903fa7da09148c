"""Sparse dynamic all-to-all ("crystal router") across ranks.

Same contract as the reference's `swirl_fem/communication/crystal_router.py:
36-115`: rank p holds `n` valid items `data[:n]` with destinations
`target[:n]`;

  n_out, data_out, source = crystal_router(n, data, target)

delivers every item to its target rank (order within a rank unspecified) and
tells the receiver where each item came from, so that a second call
`crystal_router(n_out, data_out, source)` returns the items to their senders.
`data` may be a pytree of tensors with the same leading length.

The reference routes through a hypercube of `lax.pshuffle` stages with static,
padded buffers because `pmap` collectives cannot have data-dependent sizes.
One process per GPU has no such limit: the items are bucketed by target (a
stable sort), the bucket sizes are exchanged with one small `all_to_all_single`
and the payload with ONE variable-split `all_to_all_single` per leaf (NCCL
grouped send/recv over NVLink).  Used at set-up time only (discovery of shared
dofs on arbitrary partitions, global numbering), never inside the solver loop.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def _leaves(tree):
  if isinstance(tree, dict):
    return [l for k in tree for l in _leaves(tree[k])]
  if isinstance(tree, (list, tuple)):
    return [l for t in tree for l in _leaves(t)]
  return [tree]


def _rebuild(tree, leaves):
  it = iter(leaves)

  def go(t):
    if isinstance(t, dict):
      return {k: go(v) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
      return type(t)(go(v) for v in t)
    return next(it)
  return go(tree)


def crystal_router_setup(group=None):
  """Returns the router bound to a process group (reference:
  `crystal_router_setup(mesh, axis_name)`, crystal_router.py:36)."""

  def crystal_router(n, data, target, return_source: bool = True):
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = int(n)
    leaves = _leaves(data)
    device = target.device
    for leaf in leaves:
      if leaf.shape[0] != target.shape[0]:
        raise ValueError('all leaves of `data` must have the static length of '
                         f'`target` ({target.shape[0]}), got {leaf.shape[0]}')
    if not 0 <= n <= target.shape[0]:
      raise ValueError(f'dynamic length {n} exceeds the static length '
                       f'{target.shape[0]}')
    tgt = target[:n].to(torch.int64)
    if n and (int(tgt.min()) < 0 or int(tgt.max()) >= world):
      raise ValueError('targets must be ranks in [0, world)')
    order = torch.argsort(tgt, stable=True)
    send_counts = torch.bincount(tgt, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    in_splits = send_counts.tolist()
    out_splits = recv_counts.tolist()
    n_out = int(sum(out_splits))

    def route(leaf):
      payload = leaf[:n][order].contiguous()
      out = torch.empty((n_out,) + tuple(leaf.shape[1:]), dtype=leaf.dtype,
                        device=device)
      dist.all_to_all_single(out, payload, output_split_sizes=out_splits,
                             input_split_sizes=in_splits, group=group)
      return out

    routed = [route(leaf) for leaf in leaves]
    data_out = _rebuild(data, routed)
    if not return_source:
      return n_out, data_out
    source = torch.repeat_interleave(
        torch.arange(world, device=device),
        torch.as_tensor(out_splits, device=device)).to(target.dtype)
    del rank
    return n_out, data_out, source

  return crystal_router
