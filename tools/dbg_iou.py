import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from rec_attend_b200 import ops
from oracle import model as OM
for (B,T,H,W) in [(7,20,32,64),(2,32,64,64),(6,20,32,64),(1,20,32,64)]:
  rng = np.random.default_rng(B * 100 + T)
  a = rng.random((B, T, H, W)).astype(np.float32) ** 3
  g = (rng.random((B, T, H, W)) > 0.7).astype(np.float32)
  res = ops.f_iou_soft_hard(torch.from_numpy(a).cuda(), torch.from_numpy(g).cuda())
  g[:, -1] = 0.0
  res = ops.f_iou_soft_hard(torch.from_numpy(a).cuda(), torch.from_numpy(g).cuda())
  soft = res[0].cpu().numpy()
  ah=(torch.from_numpy(a)>0.5).float()
  rh=OM.f_iou_pairwise(ah, torch.from_numpy(g)).numpy(); rd=OM.f_dice_pairwise(ah, torch.from_numpy(g)).numpy()
  print(' hard err', np.abs(res[1].cpu().numpy()-rh).max(), rh.max(), 'dice err', np.abs(res[2].cpu().numpy()-rd).max(), rd.max())
  old=ops.f_iou(torch.from_numpy(a).cuda(), torch.from_numpy(g).cuda()).cpu().numpy(); print(' vs old', np.abs(old-soft).max())
  ref = OM.f_iou_pairwise(torch.from_numpy(a), torch.from_numpy(g)).numpy()
  err = np.abs(soft-ref).max(axis=(1,2))
  print((B,T,H,W), 'per-example max err', np.round(err,5))
  b = int(err.argmax())
  d = np.abs(soft[b]-ref[b])
  print(' worst example', b, 'rows with err', np.nonzero(d.max(axis=1)>1e-4)[0][:30], 'cols', np.nonzero(d.max(axis=0)>1e-4)[0][:30])
  print(' sample', soft[b][:2,:4], ref[b][:2,:4])
