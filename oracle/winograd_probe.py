"""Numerics probe (TEST INFRASTRUCTURE / experiment record, not product code): would Winograd F(2x2, 3x3) fit the
2e-2 logit bound with bf16 operands?

    python -m oracle.winograd_probe

CPU emulation of the device data path (oracle/bf16_emulation.py conventions: BN folded in fp32, weights rounded once to
bf16, bf16 activations between layers, fp32 accumulation) with every 3x3 / stride-1 convolution of the ResNet-18 trunk
replaced by Winograd F(2x2, 3x3): input tiles transformed in fp32 (B^T d B) and rounded to bf16, weights transformed in
fp32 (G g G^T) and rounded to bf16, 16 element-wise products accumulated over channels in fp32, output transform
(A^T m A) in fp32.  Compared with the direct-convolution emulation against the LIVE reference's logits
(tests/golden/decisions_n2.npz, v2 fixture, 6 held-out segments).

Result in the build container (2026-10-18): direct 0.0031 max / 0.00087 mean, Winograd 0.0025 max / 0.00087 mean --
the transform costs no accuracy on this network, so the 2.25x MAC reduction on 16 of the 20 convolutions (87% of the
FLOPs) is open to a power-bound step.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import bf16_emulation as E, fixtures as FX, restatement as R   # noqa: E402

torch.set_num_threads(8)
q=lambda t: t.to(torch.bfloat16).float()
Bt=torch.tensor([[1,0,-1,0],[0,1,1,0],[0,-1,1,0],[0,1,0,-1]],dtype=torch.float32)
G=torch.tensor([[1,0,0],[.5,.5,.5],[.5,-.5,.5],[0,0,1]],dtype=torch.float32)
At=torch.tensor([[1,1,1,0],[0,1,-1,-1]],dtype=torch.float32)
def wino(x, w, b):
    """x [B,C,H,W] bf16-valued, w [O,C,3,3] fp32 (folded), F(2x2,3x3); transformed operands rounded to bf16, fp32 accumulate"""
    Bn,C,H,W=x.shape; O=w.shape[0]
    U=q(torch.einsum('ij,ocjk,lk->ocil',G,w,G))                 # [O,C,4,4]
    xp=F.pad(x,(1,1,1,1))
    t=xp.unfold(2,4,2).unfold(3,4,2)                            # [B,C,H/2,W/2,4,4]
    V=q(torch.einsum('ij,bchwjk,lk->bchwil',Bt,t,Bt))           # [B,C,h,w,4,4]
    M=torch.einsum('ocil,bchwil->bohwil',U,V)                   # fp32
    Y=torch.einsum('ij,bohwjk,lk->bohwil',At,M,At)              # [B,O,h,w,2,2]
    Y=Y.permute(0,1,2,4,3,5).reshape(Bn,O,H,W)
    return Y+b.view(1,-1,1,1)
def backbone(img1, sd, p, use_wino):
    w,b=E.fold_bn(sd[p+"conv1.weight"],sd,p+"bn1"); w1=q(w.sum(dim=1,keepdim=True))
    x=F.conv2d(q(img1),w1,b,stride=2,padding=3); x=q(F.relu(x)); x=F.max_pool2d(x,3,2,1)
    for li,stride in ((1,1),(2,2),(3,2),(4,2)):
        for blk in range(2):
            qn=f"{p}layer{li}.{blk}"; s=stride if blk==0 else 1
            w,b=E.fold_bn(sd[qn+".conv1.weight"],sd,qn+".bn1")
            o = wino(x,w,b) if (use_wino and s==1) else F.conv2d(x,q(w),b,stride=s,padding=1)
            o=q(F.relu(o))
            if (qn+".downsample.0.weight") in sd:
                wd,bd=E.fold_bn(sd[qn+".downsample.0.weight"],sd,qn+".downsample.1"); idn=q(F.conv2d(x,q(wd),bd,stride=s))
            else: idn=x
            w,b=E.fold_bn(sd[qn+".conv2.weight"],sd,qn+".bn2")
            o2 = wino(o,w,b) if use_wino else F.conv2d(o,q(w),b,padding=1)
            x=q(F.relu(o2+idn))
    return x
N=2; sd=FX.decision_state_dict(N)
g=np.load(os.path.join(ROOT, 'tests', 'golden', 'decisions_n2.npz'))
x,_=FX.family_segments(6,int(g['first']),n_classes=N+1)
img=R.waveform_to_image(x).unsqueeze(1)
with torch.no_grad():
    for uw in (False,True):
        t=time.time(); outs=[]
        for i in range(N):
            p=f"sub_models.{i}."
            outs.append(E.head_fp32(backbone(img,sd,p+"base.",uw),sd,p+"head."))
        z=R.merge_logits(torch.stack(outs,dim=1)).numpy()
        d=np.abs(z-g['merged_logits'][:6])
        print('winograd' if uw else 'direct  ', 'max |logit diff| vs reference', d.max(), 'mean', d.mean(), round(time.time()-t,1),'s')
