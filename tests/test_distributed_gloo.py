"""World-size-2 gloo test (CPU) of the multi-rank path: contiguous shape sharding, per-rank seeds,
ragged all-gather of finished clouds, all-reduce of metric partials."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from bdm_b200 import distributed as D
    lo, hi = D.shard_range(total)
    g = torch.Generator().manual_seed(D.rank_seed(100))
    # stand-in for a sampled shard: values identify (global shape index) so the gather order is checkable
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 3).contiguous()
    local = local + 0.0 * torch.randn(local.shape, generator=g)
    full = D.gather_samples(local, total)
    cd = torch.arange(lo, hi, dtype=torch.float64)
    f1 = torch.ones(hi - lo, dtype=torch.float64) * (rank + 1)
    mean_cd, mean_f1, n = D.reduce_metrics(cd, f1)
    torch.save(dict(full=full, mean_cd=mean_cd, mean_f1=mean_f1, n=n, lo=lo, hi=hi), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_sharding_gather_reduce_world2(tmp_path):
    total, ws = 5, 2  # ragged: 3 + 2
    mp.spawn(_worker, args=(ws, _free_port(), total, str(tmp_path)), nprocs=ws, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(ws)]
    assert (res[0]["lo"], res[0]["hi"], res[1]["lo"], res[1]["hi"]) == (0, 3, 3, 5)
    for r in res:
        assert r["full"].shape == (total, 4, 3)
        assert torch.equal(r["full"][:, 0, 0], torch.arange(total, dtype=torch.float32))
        assert r["n"] == total
        assert abs(r["mean_cd"] - 2.0) < 1e-12                  # mean of 0..4
        assert abs(r["mean_f1"] - (3 * 1 + 2 * 2) / 5) < 1e-12


def test_shard_range_partitions():
    from bdm_b200.distributed import shard_range
    for total in (0, 1, 7, 256):
        for ws in (1, 2, 4, 8):
            spans = [shard_range(total, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
