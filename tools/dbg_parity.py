import sys; sys.path.insert(0,'.')
import numpy as np, torch
import rec_attend_b200 as ra
from rec_attend_b200.full_model import FullModel
from oracle import model as OM
opt=ra.config.full_model_opt('cvppp',128,128,8); b=ra.synthetic.make_batch(opt,1); w=ra.synthetic.make_weights(opt)
ref=OM.full_model_forward(opt,w,b)
out=FullModel(opt).load_weights(w).forward(b)
for k in ['ctrl_out','attn_ctr','attn_size','attn_lg_var','attn_box','y_out','x_patch','y_out_patch','s_out','canvas','h_ctrl']:
    a=out[k].cpu().numpy(); r=ref[k].numpy()
    err=np.abs(a-r); 
    print(k, 'max rel %.2e'%(err.max()/np.abs(r).max()), 'per-step', ['%.1e'%(np.abs(a[:,t]-r[:,t]).max()/np.abs(r).max()) for t in range(8)] if a.ndim>1 and a.shape[1]==8 else '')
print('ctr', ref['attn_ctr'][0].numpy().T, 'box gamma', np.exp(ref['ctrl_out'][0,:,7].numpy()))
