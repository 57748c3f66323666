"""K-batch sharding across GPUs: every (m, k) column is an independent
transform and k is the slowest index, so rank r of `world` owns the contiguous slab
[k0, k1) at element offset k0 * stride[dim+1].  No collective is on the data path."""


def shard_k(K, world, rank, pair_align=False):
    """Contiguous slab [k0, k1) of rank `rank`; remainders go to the first ranks.  With
    pair_align the slab boundaries are even (odd-N real transforms pair rows 2k', 2k'+1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    unit = 2 if pair_align else 1
    units = (K + unit - 1) // unit
    base, rem = divmod(units, world)
    u0 = rank * base + min(rank, rem)
    u1 = u0 + base + (1 if rank < rem else 0)
    return min(u0 * unit, K), min(u1 * unit, K)


def shard_config(pkg, cfg, world, rank):
    """Per-rank configuration (same kernel, K_local slices) and the element offsets of the slab
    in the input and output tensors."""
    dim = cfg.dim
    K = cfg.shape[dim + 1]
    odd_real = cfg.type != 0 and cfg.shape[1] % 2 == 1
    k0, k1 = shard_k(K, world, rank, pair_align=odd_real)
    shape = list(cfg.shape)
    shape[dim + 1] = k1 - k0
    local = pkg.make_config(dim, shape[: dim + 2], cfg.fp, cfg.dir, cfg.type, istride=list(cfg.istride)[: dim + 2],
                            ostride=list(cfg.ostride)[: dim + 2])
    return local, k0 * cfg.istride[dim + 1], k0 * cfg.ostride[dim + 1], (k0, k1)
