"""N > 1 host logic on CPU (gloo, world_size 2): the one exchange step of the path -- the packed all-gather of
contrastive targets (pretrain/pretrain_model.py:284-290) -- must reproduce the reference's rank-major
`all_gather(y, 'batch').reshape(-1, H)`, and the sharded loss with gathered negatives must equal the
single-device loss on the global batch (SURVEY.md §8c self-check v)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.reserve_oracle import contrastive_loss


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_preds(rank, b=2, H=32):
    g = torch.Generator().manual_seed(100 + rank)
    unit = lambda n: torch.nn.functional.normalize(torch.randn(n, H, generator=g), dim=-1) * 4.0
    return {
        "imgs_to_audio": {"x": unit(16 * b), "y": unit(16 * b)},
        "text_to_audio": {"x": unit(12 * b), "y": unit(12 * b), "y_extra": unit(36 * b)},
        "stuff_to_span": {"x": unit(48 * b), "y": unit(48 * b), "_sources": torch.randint(0, 3, (48 * b,), generator=g)},
    }


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from merlot_reserve_b200.pretrain_model import gather_contrastive_targets
    preds = _make_preds(rank)
    jobs = gather_contrastive_targets(preds)
    # reference semantics: y_all = all_gather(cat(y, y_extra)).reshape(-1, H), rank-major
    expect = []
    for name in preds:
        for k1, k2 in (("x", "y"), ("y", "x")):
            per_rank = []
            for r in range(world):
                d = _make_preds(r)[name]
                per_rank.append(torch.cat([d[k2]] + ([d[k2 + "_extra"]] if (k2 + "_extra") in d else []), 0))
            expect.append(torch.cat(per_rank, 0))
    ok = len(jobs) == len(expect) and all(torch.equal(j[3], e) for j, e in zip(jobs, expect))
    # sharded loss with the gathered targets (oracle arithmetic on this rank's x rows)
    it = iter(jobs)
    loss, _ = contrastive_loss(preds, gather=lambda y: next(it)[3])
    q.put((rank, ok, float(loss)))
    dist.destroy_process_group()


def test_packed_allgather_world2_matches_reference_order_and_global_loss():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    # single-device loss on the global batch: rows of both ranks concatenated (extras stay targets only)
    all_preds = [_make_preds(r) for r in range(world)]
    glob = {}
    for name in all_preds[0]:
        glob[name] = {k: torch.cat([p[name][k] for p in all_preds], 0) for k in all_preds[0][name]}
    ref_loss, _ = contrastive_loss(glob)
    mean_sharded = sum(l for _, _, l in res) / world
    assert abs(mean_sharded - float(ref_loss)) < 1e-4


def test_gather_is_identity_without_process_group():
    from merlot_reserve_b200.pretrain_model import gather_contrastive_targets
    preds = _make_preds(0)
    jobs = gather_contrastive_targets(preds)
    assert [j[0] for j in jobs] == ["imgs_to_audio"] * 2 + ["text_to_audio"] * 2 + ["stuff_to_span"] * 2
    assert torch.equal(jobs[2][3], torch.cat([preds["text_to_audio"]["y"], preds["text_to_audio"]["y_extra"]], 0))
    assert torch.equal(jobs[3][3], preds["text_to_audio"]["x"])
