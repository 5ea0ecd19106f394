"""Read-sharded multi-GPU plumbing: the index is replicated, reads are
split into contiguous per-rank ranges, and the ONLY exchange of the path is one sum-reduce of
the counter block (plus the per-leaf rcount arrays in mode P) into rank 0.  One process per
GPU over torch.distributed (NCCL on GPUs; the same code runs over gloo on CPU tensors in the
tests).  Integer sums: the combined result does not depend on the number of shards."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of rank's items; the ranges partition [0, n_items)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return n_items * rank // world, n_items * (rank + 1) // world


def combine_counters(counts, rcount_u=None, rcount_d=None, dst=0, group=None):
    """Sum-reduce the counter block (int64 view of uint64[2(G+1)+4]) and, when given, the two
    rcount arrays (int32 views of uint32) into rank `dst`, in place.  Two's-complement addition
    is bit-identical to the unsigned sums.  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    tensors = [counts] + [t for t in (rcount_u, rcount_d) if t is not None and t.numel() > 0]
    grouped = getattr(dist, "_coalescing_manager", None)
    big = tensors[1:]
    if counts.is_cuda and len(big) > 1 and grouped is not None and all(t.dtype == big[0].dtype for t in big):
        # the two per-leaf arrays (same type) go out as ONE grouped NCCL launch (ncclGroupStart/End
        # around the sums); the coalescing manager groups all-reduces of one type, which leaves the
        # totals on rank `dst` as well.  The small counter block (another type) is its own reduce.
        dist.reduce(counts, dst=dst, op=dist.ReduceOp.SUM, group=group)
        with grouped(group=group, device=counts.device):
            for t in big:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return
    for t in tensors:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)


def gather_pair_maps(pairs, dst=0, group=None):
    """SC mode: per-rank {(a, b): count} dicts merged on rank `dst` (read_cnts_b is sparse, so
    it is gathered and merged on the host rather than reduced)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(pairs)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(dict(pairs), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    merged = {}
    for d in bucket:
        for k, v in d.items():
            merged[k] = merged.get(k, 0) + v
    return merged


def unpack_counts(counts, n_genomes):
    """Counter block -> dict (cnt_u, cnt_d as lists incl. the unused slot 0, nundet, nconf, n_invalid)."""
    c = counts.detach().cpu().tolist()
    g1 = n_genomes + 1
    return {"cnt_u": c[:g1], "cnt_d": c[g1:2 * g1], "nundet": c[2 * g1], "nconf": c[2 * g1 + 1],
            "n_invalid": c[2 * g1 + 2]}
