"""Item-axis sharding of predict_rank(k) over the GPUs of one box.

One process per GPU (torchrun); every rank holds the full user side and a contiguous range of item rows.  The path
has exactly one exchange step: an all-gather of the per-shard top-k candidates (score f32, global item id i32) --
U*k*8 bytes per rank -- followed by the same deterministic merge (trk_topk_merge) on every rank, ordered by
(score desc, global id asc) so cross-shard ties resolve to the lower id like tf.nn.top_k.  torch.distributed is the
plumbing: NCCL over NVLink on GPUs, gloo on CPU for the host-logic tests."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced item range [lo, hi) of `rank`; the first n_items % world_size ranks get one extra."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_candidates(top_scores, top_items, group=None):
    """[U, k] per rank -> ([U, world, k] scores, [U, world, k] ids), identical on every rank, lists in rank order.

    Scores and ids travel in ONE collective: both are 32-bit, so they are packed into a [2, U, k] int32 buffer."""
    world = dist.get_world_size(group)
    n_users, k = top_scores.shape
    packed = torch.stack([top_scores.contiguous().view(torch.int32), top_items.contiguous()])      # [2, U, k]
    # output = the ranks' buffers concatenated along dim 0 (the layout both NCCL and gloo accept)
    gathered = torch.empty((world * 2, n_users, k), dtype=torch.int32, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed.contiguous(), group=group)
    gathered = gathered.view(world, 2, n_users, k)
    scores = gathered[:, 0].view(torch.float32).permute(1, 0, 2).contiguous()                      # [U, world, k]
    items = gathered[:, 1].permute(1, 0, 2).contiguous()
    return scores, items


def sharded_predict_top_k(model, user_features, item_features, k, group=None, to_host=True):
    """predict_rank(k) with the item axis sharded over the ranks of `group` (default: the world group).

    `item_features` is the FULL item matrix on every rank (each rank slices its own rows); returns the global top-k,
    identical on all ranks."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(item_features.shape[0], world, rank)
    local_items = item_features.tocsr()[lo:hi] if hasattr(item_features, 'tocsr') else item_features[lo:hi]
    return model.predict_top_k(user_features, local_items, k, item_id_offset=lo,
                               gather_group=group if group is not None else dist.group.WORLD, to_host=to_host)
