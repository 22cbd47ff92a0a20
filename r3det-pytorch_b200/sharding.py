"""Multi-GPU harness for the rotated-geometry path: one process per GPU (torch.distributed), no collective
inside any op.  The path shards by independent units:
  * pairwise IoU — the anchor axis is split into contiguous row blocks, the (<= ~1k) GT boxes are replicated;
    the IoU matrix stays sharded.  Only the assigner's per-GT statistics cross ranks: one all_reduce(MAX) of
    a packed (iou_bits << 32 | ~anchor_index) int64 per GT and one all_reduce(SUM) of the pos/neg counts;
  * NMS / FRM — images are independent; each rank processes its images and ONE all_gather of fixed-size
    padded records (max_per_img x 7 floats + a count) publishes the keep lists.
The compute callables are injected so that the plumbing can be exercised on CPU with the gloo backend
(tests/test_sharding.py); in production they are the CUDA ops of this package."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n units for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def sharded_pairwise_iou(gt, anchors, iou_fn, group=None):
    """Each rank computes overlaps(gt, anchors[lo:hi]) -> ((G, hi-lo) tensor, lo, hi).  No communication."""
    rank, world = _world(group)
    lo, hi = shard_range(anchors.size(0), rank, world)
    return iou_fn(gt, anchors[lo:hi]), lo, hi


def assigner_stats(local_overlaps, lo, pos_iou_thr, neg_iou_thr, group=None):
    """Global per-GT best anchor and pos/neg anchor counts from row-sharded overlaps.

    local_overlaps: (G, n_local) IoUs of this rank's anchors [lo, lo + n_local).
    Returns (gt_max (G,), gt_argmax (G,) global anchor index, num_pos, num_neg) identical on every rank.
    Ties resolve to the lowest anchor index, like torch.max over the unsharded matrix."""
    G, n_local = local_overlaps.shape
    dev = local_overlaps.device
    if n_local > 0 and G > 0:
        vals, idx = local_overlaps.max(dim=1)
        bits = vals.clamp_min(0).contiguous().view(torch.int32).to(torch.int64)          # IoU >= 0: bit order == value order
        packed = (bits << 32) | (0xFFFFFFFF - (idx.to(torch.int64) + lo))
        amax = local_overlaps.max(dim=0)[0]
        counts = torch.stack([(amax >= pos_iou_thr).sum(), ((amax >= 0) & (amax < neg_iou_thr)).sum()]).to(torch.int64)
    else:
        packed = torch.zeros((G,), dtype=torch.int64, device=dev)
        counts = torch.zeros((2,), dtype=torch.int64, device=dev)
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    gt_max = (packed >> 32).to(torch.int32).view(torch.float32)
    gt_argmax = 0xFFFFFFFF - (packed & 0xFFFFFFFF)
    return gt_max, gt_argmax, int(counts[0]), int(counts[1])


def gather_keep_lists(local_dets, local_labels, max_per_img, num_images, group=None):
    """Publish per-image detections with ONE all_gather.

    local_dets: list over this rank's images (shard_range(num_images, rank, world)) of (k_i, 6) tensors,
    local_labels: matching (k_i,) int64 tensors, k_i <= max_per_img.
    Returns lists (dets, labels) over all `num_images` images, identical on every rank."""
    rank, world = _world(group)
    per_rank = (num_images + world - 1) // world
    dev = local_dets[0].device if local_dets else torch.device('cpu')
    rec = torch.zeros((per_rank, max_per_img, 7), dtype=torch.float32, device=dev)
    cnt = torch.zeros((per_rank,), dtype=torch.float32, device=dev)
    for i, (d, l) in enumerate(zip(local_dets, local_labels)):
        k = min(d.size(0), max_per_img)
        rec[i, :k, :6] = d[:k]
        rec[i, :k, 6] = l[:k].to(torch.float32)
        cnt[i] = k
    payload = torch.cat([rec.reshape(per_rank, -1), cnt[:, None]], 1).contiguous()   # one fixed-size message per rank
    if world > 1:
        bufs = [torch.empty_like(payload) for _ in range(world)]
        dist.all_gather(bufs, payload, group=group)
    else:
        bufs = [payload]
    dets, labels = [], []
    for r in range(world):
        lo, hi = shard_range(num_images, r, world)
        for i in range(hi - lo):
            k = int(bufs[r][i, -1].item())
            block = bufs[r][i, :-1].reshape(max_per_img, 7)[:k]
            dets.append(block[:, :6].clone())
            labels.append(block[:, 6].to(torch.int64))
    return dets, labels
