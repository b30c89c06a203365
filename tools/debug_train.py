import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cases
from tests.test_gpu_train import _inputs, _make_trainer, rel_err
from oracle import fusion_train as ot
B = int(os.environ.get("TB", "6"))
model, sd, tr = _make_trainer(0.0)
g, img, txt, labels = _inputs(B=B)
hb = cases.to_host_batch(g)
tr.debug_taps = {}
loss, logits = tr.forward_backward(g.to("cuda"), img.cuda(), txt.cuda(), labels.cuda())
torch.cuda.synchronize()
T = {k: v.float().cpu() for k, v in tr.debug_taps.items()}
grads = {k: v.cpu().clone() for k, v in tr.named_grads().items()}
# A. fp32 oracle end to end
taps = {}
l0, lg0, g0 = ot.loss_and_grads(sd, hb, img, txt, labels, taps=taps)
print("A  gcn_in", rel_err(T["gcn_in"].view(B, 100, 512), taps["gcn_in"]), "loss", float(loss), float(l0))
# B. relay: oracle chain from the CUDA gcn_in
taps = {}
l1, lg1, g1 = ot.loss_and_grads(sd, hb, img, txt, labels, taps=taps, gcn_in=T["gcn_in"].view(B, 100, 512), emulate_bf16=True)
f = tr.last["feats"].cpu()
print("B  gcn_k", [round(rel_err(T[f"gcn_{k}"].view(B, 100, 512), taps[f"gcn_{k}"]), 6) for k in range(1, 9)])
print("B  feats", [round(rel_err(f[:, lo:lo+512], taps["feats"][:, lo:lo+512]), 6) for lo in (0, 512, 1024)])
print("B  loss", float(loss), float(l1), "logits err/scale", float((logits.cpu() - lg1).abs().max() / lg1.abs().max()))
print("C  d_gcn_in", rel_err(T["d_gcn_in"].view(B, 100, 512), g1["__gcn_in__"]))
names = [n for n in g1 if n != "__gcn_in__"]
worst = sorted(((rel_err(grads[n].reshape(g1[n].shape), g1[n]), n, float(g1[n].norm())) for n in names), reverse=True)
for w in worst[:12]: print("   ", w)
flat = torch.cat([grads[n].reshape(-1) for n in names]); flat_ref = torch.cat([g1[n].reshape(-1) for n in names])
print("C  downstream flat grad err", rel_err(flat, flat_ref), "median", worst[len(worst)//2])
# D. relay backward: graph branch grads from the CUDA cotangent
g2 = ot.graph_branch_grads(sd, hb, T["d_gcn_in"].view(B, 100, 512), emulate_bf16=True)
names = list(g2)
worst = sorted(((rel_err(grads[n].reshape(g2[n].shape), g2[n]), n, float(g2[n].norm())) for n in names), reverse=True)
for w in worst[:12]: print("   ", w)
flat = torch.cat([grads[n].reshape(-1) for n in names]); flat_ref = torch.cat([g2[n].reshape(-1) for n in names])
print("D  upstream flat grad err", rel_err(flat, flat_ref), "median", worst[len(worst)//2])
