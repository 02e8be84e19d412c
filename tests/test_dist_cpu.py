"""World-size-2 gloo test (CPU) of the data-parallel host logic: the flat gradient arena is summed over ranks by
`aclgan_Trainer._allreduce` and the 1/world scale the Adam kernel applies reproduces "N reference replicas at local
batch B with averaged gradients" (SURVEY.md 8e).  The per-rank gradients come from the CPU oracle here (the CUDA
kernels need a GPU); the semantics under test are the arena / all-reduce / scaling plumbing."""
import copy
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg():
    cfg = yaml.safe_load(open(os.path.join(ROOT, "acl-gan_b200", "configs", "selfie2anime.yaml")))
    cfg["gen"].update(dim=8, mlp_dim=16, n_res=1)
    cfg["dis"].update(dim=8, n_layer=3)
    cfg["display_size"] = 1
    return cfg


def _rank_grads(rank, cfg):
    import aclgan_oracle as O
    torch.manual_seed(0)                      # identical weights on every rank
    ot = O.OracleTrainer(copy.deepcopy(cfg), dtype=torch.float64)
    g = torch.Generator().manual_seed(1234 + rank)
    x_a = (torch.rand(1, 3, 32, 32, generator=g) * 2 - 1).double()
    x_b = (torch.rand(1, 3, 32, 32, generator=g) * 2 - 1).double()
    zs = [torch.randn(1, 8, 1, 1, generator=g).double() for _ in range(3)]
    ot.dis_update(x_a, x_b, zs, step=False)
    return torch.cat([v.grad.reshape(-1) for n in ("dis_A", "dis_B", "dis_2") for v in ot.nets[n].values()
                      if v.grad is not None])


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "acl-gan_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import engine as E
    import trainer as T
    cfg = _cfg()
    flat = _rank_grads(rank, cfg)

    class Arena:                # duck-typed engine.GradArena on the CPU
        pass
    arena = Arena()
    arena.flat = flat.clone()
    T.aclgan_Trainer._allreduce(None, arena)             # the product code path (NCCL on GPUs, gloo here)
    scaled = arena.flat / world                          # == grad_scale operand of the Adam kernel (hyper[5])
    if rank == 0:
        torch.save(dict(avg=scaled, mine=flat), out)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_averages_replica_gradients(tmp_path):
    world, port = 2, _free_port()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = torch.load(out)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cfg = _cfg()
    expect = (_rank_grads(0, cfg) + _rank_grads(1, cfg)) / 2
    assert torch.allclose(res["avg"], expect, rtol=1e-12, atol=1e-14)
    assert not torch.allclose(res["mine"], expect)       # the ranks really saw different shards
