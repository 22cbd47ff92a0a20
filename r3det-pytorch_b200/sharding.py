"""Multi-GPU harness for the rotated-geometry path: one process per GPU (torch.distributed), no collective
inside any op.  The path shards by independent units:
  * assignment — the anchor axis is split into contiguous row blocks, the (<= ~1k) GT boxes are replicated.  Every rank runs
    the FUSED assigner on its block (no (G, A / R) overlap matrix exists anywhere); only the per-GT statistics cross
    ranks: ONE all_gather of G packed (iou_bits << 32 | ~anchor_index) words + the two pos / neg counters per rank;
  * NMS / FRM — images are independent; each rank processes its images and ONE all_gather of fixed-size
    padded records (max_per_img x 7 floats + a count) publishes the keep lists; it can be issued asynchronously so that
    it overlaps the NMS of the next batch (NCCL runs it on its own stream).
The compute callables are injected so that the plumbing can be exercised on CPU with the gloo backend
(tests/test_sharding.py); in production they are the CUDA ops of this package (reference call sites:
r3det/models/dense_heads/rotate_anchor_head.py:220-228, 316-333 for the assignment, :626-673 for the keep lists)."""
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


def sharded_assign(gt, anchors, assign_fn, group=None):
    """Each rank assigns its anchor block: assign_fn(gt, anchors[lo:hi]) -> an object with `max_overlaps` (n_local,),
    `gt_max_overlaps` (G,), `gt_argmax_overlaps` (G,) LOCAL anchor indices (r3det_b200.max_iou_assign's AssignOutput).
    Returns (output, lo, hi).  No communication."""
    rank, world = _world(group)
    lo, hi = shard_range(anchors.size(0), rank, world)
    return assign_fn(gt, anchors[lo:hi]), lo, hi


def assigner_stats(gt_max_local, gt_argmax_local, max_overlaps_local, lo, pos_iou_thr, neg_iou_thr, group=None, sync=True):
    """Global per-GT best anchor and pos / neg anchor counts from the row-sharded fused assignment.

    gt_max_local / gt_argmax_local: (G,) best overlap of each GT inside this rank's anchors [lo, lo + n_local) and its LOCAL
    index; max_overlaps_local: (n_local,) best overlap of each local anchor.  One all_gather of G + 2 int64 per rank.
    Returns (gt_max (G,), gt_argmax (G,) global anchor index, num_pos, num_neg), identical on every rank; ties resolve to
    the lowest anchor index, like torch.max over the unsharded matrix.  sync=False returns the two counts as 0-dim device
    tensors instead of Python ints (no host read at all)."""
    G = gt_max_local.numel()
    dev = gt_max_local.device
    bits = gt_max_local.float().clamp_min(0).contiguous().view(torch.int32).to(torch.int64)    # IoU >= 0: bit order == value order
    packed = (bits << 32) | (0xFFFFFFFF - (gt_argmax_local.to(torch.int64) + lo))
    amax = max_overlaps_local
    counts = torch.stack([(amax >= pos_iou_thr).sum(), ((amax >= 0) & (amax < neg_iou_thr)).sum()]).to(torch.int64)
    msg = torch.cat([packed, counts.to(dev)])
    rank, world = _world(group)
    if world > 1:
        flat = torch.empty((world * (G + 2),), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(flat, msg, group=group)
        buf = flat.view(world, G + 2)
    else:
        buf = msg[None]
    best = buf[:, :G].max(dim=0).values
    tot = buf[:, G:].sum(dim=0)
    gt_max = (best >> 32).to(torch.int32).view(torch.float32)
    gt_argmax = 0xFFFFFFFF - (best & 0xFFFFFFFF)
    if not sync:
        return gt_max, gt_argmax, tot[0], tot[1]
    npos, nneg = tot.tolist()                                                                   # the only host read
    return gt_max, gt_argmax, int(npos), int(nneg)


def pack_keep_records(local_dets, local_labels, max_per_img, per_rank):
    """(per_rank, max_per_img * 7 + 1) float32 payload: rows <x, y, w, h, a, score, label> padded to max_per_img + the count."""
    dev = local_dets[0].device if local_dets else torch.device('cpu')
    rec = torch.zeros((per_rank, max_per_img, 7), dtype=torch.float32, device=dev)
    cnt = torch.zeros((per_rank,), dtype=torch.float32, device=dev)
    for i, (d, l) in enumerate(zip(local_dets, local_labels)):
        k = min(d.size(0), max_per_img)
        rec[i, :k, :6] = d[:k]
        rec[i, :k, 6] = l[:k].to(torch.float32)
        cnt[i] = k
    return torch.cat([rec.reshape(per_rank, -1), cnt[:, None]], 1).contiguous()    # one fixed-size message per rank


class KeepListGather(object):
    """Handle of an all_gather of padded keep records in flight; `wait()` returns (dets, labels) lists over all images."""

    def __init__(self, payload, max_per_img, num_images, group=None, async_op=False):
        self.rank, self.world = _world(group)
        self.max_per_img, self.num_images = max_per_img, num_images
        self.work = None
        if self.world > 1:
            flat = torch.empty((self.world * payload.numel(),), dtype=payload.dtype, device=payload.device)
            self.work = dist.all_gather_into_tensor(flat, payload.reshape(-1), group=group, async_op=async_op)
            self.buf = flat.view((self.world,) + tuple(payload.shape))
        else:
            self.buf = payload[None]

    def wait_padded(self):
        """(dets (images, max_per_img, 6), labels (images, max_per_img) int64, counts (images,) int64) on the device, images in
        global order — no host read, no per-image work (ranks hold equal shares when num_images divides by the world size;
        otherwise the trailing ranks' unused slots are dropped)."""
        if self.work is not None:
            self.work.wait()
            self.work = None
        m = self.max_per_img
        per_rank = self.buf.size(1)
        rows = []
        for r in range(self.world):
            lo, hi = shard_range(self.num_images, r, self.world)
            rows.append(self.buf[r, :hi - lo])
        flat = torch.cat(rows) if len(rows) > 1 else rows[0]
        rec = flat[:, :-1].reshape(-1, m, 7)
        return rec[..., :6], rec[..., 6].to(torch.int64), flat[:, -1].to(torch.int64)

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        m = self.max_per_img
        counts = self.buf[:, :, -1].to(torch.int64).cpu().tolist()           # ONE device-to-host read for the whole batch
        dets, labels = [], []
        for r in range(self.world):
            lo, hi = shard_range(self.num_images, r, self.world)
            for i in range(hi - lo):
                block = self.buf[r, i, :-1].reshape(m, 7)[:counts[r][i]]
                dets.append(block[:, :6].clone())
                labels.append(block[:, 6].to(torch.int64))
        return dets, labels


def gather_keep_lists(local_dets, local_labels, max_per_img, num_images, group=None, async_op=False):
    """Publish per-image detections with ONE all_gather.

    local_dets: list over this rank's images (shard_range(num_images, rank, world)) of (k_i, 6) tensors,
    local_labels: matching (k_i,) int64 tensors, k_i <= max_per_img.
    Returns lists (dets, labels) over all `num_images` images, identical on every rank — or, with async_op=True, a
    KeepListGather whose wait() returns them (the collective then overlaps whatever is enqueued next)."""
    rank, world = _world(group)
    per_rank = (num_images + world - 1) // world
    h = KeepListGather(pack_keep_records(local_dets, local_labels, max_per_img, per_rank), max_per_img, num_images, group, async_op)
    return h if async_op else h.wait()


def gather_padded_records(dets, labels, counts, num_images, group=None, async_op=False, padded=False):
    """The same exchange from the padded device outputs of the batched NMS (no per-image Python work, no host read before the
    collective): dets (per_rank, max_per_img, 6), labels (per_rank, max_per_img) int64, counts (per_rank,) int64.
    padded=True returns the gathered records as padded device tensors (KeepListGather.wait_padded) instead of per-image lists:
    the whole exchange then runs without a single host synchronisation."""
    per_rank, m = dets.size(0), dets.size(1)
    payload = torch.cat([torch.cat([dets, labels.to(torch.float32)[..., None]], -1).reshape(per_rank, -1),
                         counts.to(torch.float32)[:, None]], 1).contiguous()
    h = KeepListGather(payload, m, num_images, group, async_op)
    if async_op:
        return h
    return h.wait_padded() if padded else h.wait()
