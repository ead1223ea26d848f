"""N = 2 on real GPUs (skipped with fewer than two): the packed NCCL all-gather of contrastive targets inside
loss_fn_given_preds (pretrain/pretrain_model.py:284-295), one process per GPU.  The mean over ranks of the
per-rank loss must equal the oracle's single-device loss on the global batch."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_preds(rank, b=4, H=768):
    g = torch.Generator().manual_seed(100 + rank)
    unit = lambda n: (torch.nn.functional.normalize(torch.randn(n, H, generator=g), dim=-1) * 4.47).to(torch.bfloat16)
    return {
        "imgs_to_audio": {"x": unit(16 * b), "y": unit(16 * b)},
        "text_to_audio": {"x": unit(12 * b), "y": unit(12 * b), "y_extra": unit(36 * b)},
        "stuff_to_span": {"x": unit(48 * b), "y": unit(48 * b),
                          "_sources": torch.randint(0, 3, (48 * b,), generator=g).to(torch.int32)},
    }


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from merlot_reserve_b200 import loss_fn_given_preds
    preds = {h: {k: v.cuda() for k, v in d.items()} for h, d in _make_preds(rank).items()}
    loss, info = loss_fn_given_preds(preds)
    torch.cuda.synchronize()
    q.put((rank, float(loss), {k: float(v) for k, v in info.items()}))
    dist.destroy_process_group()


def test_nccl_packed_allgather_loss_world2(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle.reserve_oracle import contrastive_loss
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_preds = [_make_preds(r) for r in range(world)]
    glob = {h: {k: torch.cat([p[h][k] for p in all_preds], 0) for k in all_preds[0][h]} for h in all_preds[0]}
    ref_loss, ref_info = contrastive_loss({h: {k: (v.float() if k != "_sources" else v) for k, v in d.items()}
                                           for h, d in glob.items()})
    mean_loss = sum(l for _, l, _ in res) / world
    assert abs(mean_loss - float(ref_loss)) < 5e-3, (mean_loss, float(ref_loss))
    for k in ("imgs_to_audio", "text_to_audio", "stuff_to_span"):
        m = sum(info[k] for _, _, info in res) / world
        assert abs(m - float(ref_info[k])) < 3e-3, k
