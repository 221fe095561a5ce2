"""CPU, world_size 2, gloo: the N>1 data path (image sharding + the single fixed-shape all-gather + re-ordering)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT


def _worker(rank, world, port, num_images, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from mqdet_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_indices(num_images, rank, world)
    # "detections" of image i are filled with i so the re-ordering can be checked
    det = torch.stack([torch.full((128, 6), float(i)) for i in mine])
    num = torch.tensor([i + 1 for i in mine], dtype=torch.int32)
    det_all, num_all = parallel.all_gather_detections(det, num)
    d, n = parallel.unshard(det_all, num_all, num_images, world)
    ok = d.shape == (num_images, 128, 6) and all(float(d[i, 0, 0]) == i and int(n[i]) == i + 1 for i in range(num_images))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_all_gather_world2():
    world, num_images = 2, 8
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, num_images, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _chunk_worker(rank, world, port, num_chunks, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from mqdet_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []
    orig = dist.all_gather_into_tensor
    dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    mine = parallel.shard_chunks(num_chunks, rank, world)
    B, max_out = 2, 8
    local = torch.stack([parallel.pack(torch.full((B, max_out, 6), float(c)), torch.full((B,), c + 1, dtype=torch.int32))
                         for c in mine])
    allc = parallel.all_gather_chunks(local, num_chunks)
    ok = allc.shape == (num_chunks, B, max_out + 1, 6) and len(calls) == 1
    for c in range(num_chunks):
        det, num = parallel.unpack(allc[c])
        ok = ok and float(det[0, 0, 0]) == c and int(num[0]) == c + 1
    # the image-sharded path also issues exactly ONE collective
    calls.clear()
    parallel.all_gather_packed(parallel.pack(torch.zeros(B, max_out, 6), torch.ones(B, dtype=torch.int32)))
    ret[rank] = bool(ok and len(calls) == 1)
    dist.barrier()
    dist.destroy_process_group()


def test_chunk_sharding_world2_one_collective():
    """text-column (prompt-chunk) sharding: 5 chunks over 2 ranks (3 + 2), ONE all-gather, chunk order restored."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_chunk_worker, args=(world, port, 5, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_single_process_is_identity():
    sys.path.insert(0, ROOT)
    from mqdet_b200 import parallel
    det, num = torch.zeros(3, 128, 6), torch.ones(3, dtype=torch.int32)
    a, b = parallel.all_gather_detections(det, num)
    assert a is det and b is num
    assert parallel.shard_indices(10, 1, 4) == [1, 5, 9]
    d2, n2 = parallel.unpack(parallel.pack(det, num))
    assert torch.equal(d2, det) and torch.equal(n2, num)


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from mqdet_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    shapes = {"encoder.qv_layer.0.ff_gate": (1,), "encoder.qv_layer.0.attn.to_q.weight": (512, 768), "pre_select.layers.0.ff.norm.bias": (256,)}
    grads = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
    mine = {k: v.clone() for k, v in grads.items()}
    parallel.all_reduce_gradients(grads)
    other = torch.Generator().manual_seed(100 + (1 - rank))
    exp = {k: (mine[k] + torch.randn(*s, generator=other)) / 2 for k, s in shapes.items()}
    ret[rank] = all(torch.allclose(grads[k], exp[k], atol=1e-6) for k in shapes)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_all_reduce_world2_one_collective():
    """Training side: every gradient of a step in ONE flat all-reduce, averaged over the ranks (DDP semantics)."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
