for i in 1 2 3; do python -m pytest tests/test_gpu_umma.py -q -m gpu -k "dense_backward" 2>&1 | grep -E "^E|passed|failed|assert" | head -20; done
COR_B200_LIB= python - <<'PY'
import numpy as np, torch
from cor_b200 import ops, synth
Nr,Nq,D=9000,129,256
g=synth.make_gallery(31,Nr,Nq,D=D)
t=(np.arange(Nq)*11)%Nr
dev=torch.device("cuda:0")
def cu(x,grad=False):
    y=torch.from_numpy(x).to(dev)
    return y.requires_grad_(True) if grad else y
res={}
for eng in ("auto","stream","auto","auto"):
    r,q=cu(g["regions"],True),cu(g["queries"],True)
    loss=ops.infonce_loss(r,q,cu(t),tau=0.07,engine=eng)
    (3.0*loss).backward()
    torch.cuda.synchronize()
    res.setdefault(eng,[]).append((float(loss),r.grad.float().cpu(),q.grad.float().cpu()))
s=res["stream"][0]
for k,a in enumerate(res["auto"]):
    for i in (1,2):
        d=(a[i]-s[i])
        rows=(d.norm(dim=1)/ (s[i].norm(dim=1)+1e-12))
        bad=(rows>2e-2).nonzero().flatten()
        print("run",k,"tensor",i,"rel",float(d.norm()/s[i].norm()),"nbad",bad.numel(),bad[:20].tolist(), "nan", int(torch.isnan(a[i]).sum()))
PY
