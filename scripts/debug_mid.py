"""Debug: per-tensor errors of ctr_tower_mid vs float64 for one shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_tower as T
from recsys_b200 import ops

def run(B, sizes, nz, p):
    cuda = torch.device("cuda", 0)
    torch.manual_seed(B)
    shapes = {"b1": (1,), "head.w": (nz + 1, 1), "head.b": (1,), "t.out.w": (sizes[-1], 1), "t.out.b": (1,)}
    for l, (i, o) in enumerate(zip(sizes[:-1], sizes[1:])):
        shapes.update({"t.%d.w" % l: (i, o), "t.%d.b" % l: (o,), "t.%d.bn.gamma" % l: (o,),
                       "t.%d.bn.beta" % l: (o,), "t.%d.bn.mean" % l: (o,), "t.%d.bn.var" % l: (o,)})
    frozen = [n for n in shapes if n.endswith((".bn.mean", ".bn.var"))]
    dense = ops.DenseParams(shapes, cuda, frozen=frozen)
    with torch.no_grad():
        for n in dense.names:
            v = dense[n]
            if n.endswith(".bn.var"): v.copy_(torch.rand_like(v) + 0.5)
            elif n.endswith(".w"): v.copy_(torch.randn_like(v) * (2.0 / v.shape[0]) ** 0.5)
            else: v.copy_(torch.randn_like(v) * 0.3 + (1.0 if n.endswith("gamma") else 0.0))
    adam = ops.TFAdamState(device=cuda); adam.next_lr_t()
    tw = ops.FusedTower(dense, "t", sizes, True, p, adam, seed=11)
    X = torch.randn(B, sizes[0], device=cuda)
    zs = [torch.randn(B, device=cuda) for _ in range(nz)]
    labels = (torch.rand(B, device=cuda) < 0.3).float()
    masks = T._masks(tw, B, cuda) if p > 0 else None
    P64 = {n: dense[n].detach().double().requires_grad_(n not in frozen) for n in dense.names}
    X = X.contiguous().requires_grad_(True)
    zs = [z.contiguous().requires_grad_(True) for z in zs]
    X64 = X.detach().double().requires_grad_(True)
    zs64 = [z.detach().double().requires_grad_(True) for z in zs]
    loss, logits, prob = ops.tower_head(tw, X, zs, labels, training=True)
    torch.cuda.synchronize()
    gates = [(a > 0).double() for a in tw.last_acts] + [(tw.last_y > 0).double()]
    loss64, logit64 = T._tower_head_ref(P64, X64, zs64, labels.double(), sizes, masks, p, True, gates)
    loss64.backward(); loss.backward(); tw.join(); torch.cuda.synchronize()
    def rep(name, got, want):
        err = (got.double() - want).abs()
        print("%-16s max err %.3e  scale %.3e  rel %.2e  argmax %s" % (name, err.max().item(), want.abs().max().item(),
              err.max().item() / max(want.abs().max().item(), 1e-12), tuple(int(i) for i in torch.unravel_index(err.argmax(), err.shape)) if err.dim() else ()))
        return err
    print("== B %d sizes %s p %g" % (B, sizes, p))
    rep("logits", logits, logit64.detach())
    e = rep("X.grad", X.grad, X64.grad)
    rows = (e.max(1).values > 1e-4 * X64.grad.abs().max()).nonzero().flatten()
    print("bad rows:", rows[:20].tolist(), "count", rows.numel())
    for z, z64 in zip(zs, zs64): rep("z.grad", z.grad, z64.grad)
    for n in dense.names:
        if n not in frozen: rep(n, dense[n].grad, P64[n].grad.reshape(dense[n].shape))

run(4096, [624, 100, 100], 2, 0.5)
run(4096, [624, 100, 100], 2, 0.0)
run(1000, [624, 100, 100], 2, 0.5)
