import sys; sys.path.insert(0, ".")
import torch, numpy as np, r3det_b200 as R
from tests.util import rand_obb
from oracle import port
dev=torch.device("cuda:0")
for v in ("v1","v2","v3"):
    for lo,hi,seed in ((8,512,3),(2,64,4),(1,1000,5),(4,300,6)):
        a=rand_obb(300,seed,v,lo,hi); b=rand_obb(6000,seed+10,v,lo,hi)
        got=R.pairwise_iou(torch.from_numpy(a).to(dev),torch.from_numpy(b).to(dev),v,"iof").cpu().numpy()
        want=port.iou_matrix(a,b,v,"iof",wrapper_mask=False)
        err=np.abs(got-want); i,j=np.unravel_index(err.argmax(), err.shape)
        print(v,lo,hi,"max err %.2e"%err.max(), "n>1e-5:", int((err>1e-5).sum()), "worst A wh", a[i,2:4], "B wh", b[j,2:4], "vals", got[i,j], want[i,j])
