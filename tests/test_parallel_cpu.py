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


def test_single_process_is_identity():
    sys.path.insert(0, ROOT)
    from mqdet_b200 import parallel
    det, num = torch.zeros(3, 128, 6), torch.ones(3, dtype=torch.int32)
    a, b = parallel.all_gather_detections(det, num)
    assert a is det and b is num
    assert parallel.shard_indices(10, 1, 4) == [1, 5, 9]
