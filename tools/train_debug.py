import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from merlot_reserve_b200.synthetic import synthetic_params, synthetic_pretrain_batch, tiny_config
from merlot_reserve_b200.training import PretrainStep
from oracle.reserve_oracle import OracleMerlotReserve, contrastive_loss

def tt(tree):
    if isinstance(tree, dict): return {k: tt(v) for k, v in tree.items()}
    return torch.from_numpy(np.asarray(tree, dtype=np.float32)).clone().requires_grad_(True)

cfg = tiny_config(); params = synthetic_params(cfg, 0, float(sys.argv[1]) if len(sys.argv) > 1 else 4.0); batch = synthetic_pretrain_batch(cfg, 2, seed=3)
tp = tt(params)
preds = OracleMerlotReserve(cfg, tp).pretrain_forward(batch)
for h in preds.values():
    for k, v in h.items():
        if k != "_sources": v.retain_grad()
loss, _ = contrastive_loss(preds); loss.backward()
step = PretrainStep(cfg, params); step._keep_locals = True
l2, info, grads = step.loss_and_grads(batch)
d = step._dbg
def cmp(name, got, ref):
    got, ref = got.float().cpu().flatten(), ref.float().flatten()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    print(f"{name:40s} cos {cos:+.4f}  |got| {got.abs().max():.4g} |ref| {ref.abs().max():.4g}")
for name, h in preds.items():
    for k, v in h.items():
        if k != "_sources":
            cmp(f"dpred {name}/{k}", d["dpred"][name][k], v.grad)
            cmp(f" pred {name}/{k}", d["preds"][name][k], v.detach())
got = step.grads_to_flax(grads)
allg = {}
def walk(t, pre=""):
    for k, v in t.items():
        n = f"{pre}/{k}" if pre else k
        if isinstance(v, dict): walk(v, n)
        else: allg[n] = v
walk(tp)
rows = []
for n, v in allg.items():
    g = torch.from_numpy(got[n]).float().flatten(); r = v.grad.float().flatten()
    rows.append((torch.nn.functional.cosine_similarity(g, r, dim=0).item(), g.abs().max().item(), r.abs().max().item(), n))
rows.sort()
for c, a, b, n in rows[:25]: print(f"{c:+.4f} got {a:.3e} ref {b:.3e} {n}")
print("n bad", sum(1 for c, *_ in rows if c < 0.99), "of", len(rows))
for n in ["contrastive_scales", "head/kernel", "head/bias", "joint_transformer/final_ln/scale", "span_encoder/transformer/cls_proj/kernel",
          "span_encoder/transformer/final_ln/scale", "span_encoder/transformer/layer_01/mlp_layer/out/kernel", "token_encoder/Embed_0/embedding",
          "audio_encoder/transformer/cls_proj/kernel", "vision_encoder/seq_attnpool/out/kernel", "vision_encoder/embedding/kernel"]:
    node = tp
    for part in n.split("/"): node = node[part]
    cmp(n, torch.from_numpy(got[n]), node.grad)
