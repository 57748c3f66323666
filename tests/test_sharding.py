"""CPU tier: the multi-GPU path shards K with no collective; checked with 2 gloo ranks on CPU
(each rank transforms its slab with the oracle standing in for the device, then the slabs are
gathered only to verify them -- the gather is test plumbing, not part of the path)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_k_partitions():
    sh = importlib.import_module("double-batched-fft-library_b200.sharding")
    for K in (1, 7, 8, 64, 1000, 16385):
        for world in (1, 2, 3, 8):
            for pair in (False, True):
                slabs = [sh.shard_k(K, world, r, pair) for r in range(world)]
                assert slabs[0][0] == 0 and slabs[-1][1] == K
                for a, b in zip(slabs, slabs[1:]):
                    assert a[1] == b[0]
                    if pair:
                        assert a[1] % 2 == 0 or a[1] == K
                sizes = [b - a for a, b in slabs]
                assert max(sizes) - min(sizes) <= (3 if pair else 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("double-batched-fft-library_b200")
    sh = importlib.import_module("double-batched-fft-library_b200.sharding")
    from oracle import oracle
    M, N, K = 4, 12, 600  # large enough that the CTA batch size does not depend on the slab
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((K, N, M)) + 1j * rng.standard_normal((K, N, M))).astype(np.complex128)
    cfg = pkg.make_config(1, [M, N, K], 8, pkg.FORWARD, pkg.C2C, inplace=False)
    local, ioff, ooff, (k0, k1) = sh.shard_config(pkg, cfg, world, rank)
    assert ioff == k0 * M * N and ooff == k0 * M * N and local.shape[2] == k1 - k0
    # the same plan (kernel identifier) serves every slab: K is a run-time argument
    ident = pkg.describe(local)["identifier"]
    xin = np.ascontiguousarray(x.reshape(-1)[ioff: ioff + (k1 - k0) * M * N])
    out = np.empty_like(xin)
    ocfg = oracle.make_config(1, [M, N, k1 - k0], 8, -1, 0, inplace=False)
    oracle.dft(ocfg, xin, out)
    # verification only: gather the slabs on rank 0
    full = np.zeros(K * N * M, dtype=np.complex128)
    full[ooff: ooff + out.size] = out
    t = torch.from_numpy(full.view(np.float64).copy())
    dist.all_reduce(t)
    idents = [None] * world
    dist.all_gather_object(idents, ident)
    if rank == 0:
        got = t.numpy().view(np.complex128).reshape(K, N, M)
        ref = np.fft.fft(x, axis=1)
        q.put((float(np.linalg.norm(got - ref) / np.linalg.norm(ref)), len(set(idents))))
    dist.destroy_process_group()


def test_two_rank_k_sharding_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    err, n_ident = q.get(timeout=10)
    assert err < 1e-13
    assert n_ident == 1
