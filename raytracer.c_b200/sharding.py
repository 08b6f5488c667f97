"""Sample sharding across GPUs: the path is embarrassingly parallel over
samples.  Rank r of G renders the global sample indices [r*spp, (r+1)*spp) (weak scaling:
fixed work per GPU) or an even split of a fixed total (strong scaling); the Philox counter
carries the GLOBAL sample index, so the union over ranks is the same estimator as one GPU
rendering all the samples.  The per-GPU float sums are added by ONE reduce to rank 0 (NCCL
over NVLink on the GPUs, gloo in the CPU tests); rank 0 then runs the gamma/quantise kernel
with the total sample count.  There is no other collective on the data path."""


def shard_weak(rank, world, spp_per_gpu):
    """-> (sample_begin, sample_end, total_samples)"""
    if not (0 <= rank < world) or spp_per_gpu < 0:
        raise ValueError("bad shard request")
    return rank * spp_per_gpu, (rank + 1) * spp_per_gpu, world * spp_per_gpu


def shard_strong(rank, world, total_spp):
    """even split of a fixed total; the first (total % world) ranks take one extra sample"""
    if not (0 <= rank < world) or total_spp < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_spp, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end, total_spp


def reduce_to_root(accum, world):
    """the single collective: SUM of the per-rank float accumulation buffers onto rank 0"""
    if world > 1:
        import torch.distributed as dist
        dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
    return accum
