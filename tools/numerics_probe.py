"""Numerics probe (CPU): how far does an fp16-operand / fp32-accumulate LSTM stack with
fp16 activation storage drift from the fp32 reference FN_SSL?  Development tool only."""
import sys, types, torch, time
for m in ['matplotlib','matplotlib.pyplot','soundfile','webrtcvad']:
    sys.modules[m]=types.ModuleType(m)
sys.path.insert(0,'/root/reference/FN-SSL/Lightning')
import Model

def q(x, mode):
    if mode=='fp16': return x.half().float()
    if mode=='bf16': return x.bfloat16().float()
    return x

def lstm_dir(x, wih, whh, b, reverse, mode):
    # x: (N, L, I)
    N,L,I = x.shape; H = whh.shape[1]
    wih=q(wih,mode); whh=q(whh,mode)
    h=torch.zeros(N,H); c=torch.zeros(N,H); out=torch.empty(N,L,H)
    xs = q(x,mode)
    gx = xs @ wih.t() + b
    rng = range(L-1,-1,-1) if reverse else range(L)
    for t in rng:
        g = gx[:,t] + q(h,mode) @ whh.t()
        i,f,gg,o = g.chunk(4,1)
        c = torch.sigmoid(f)*c + torch.sigmoid(i)*torch.tanh(gg)
        h = torch.sigmoid(o)*torch.tanh(c)
        out[:,t]=h
    return out

def lstm(x, m, mode):
    outs=[lstm_dir(x, m.weight_ih_l0, m.weight_hh_l0, m.bias_ih_l0+m.bias_hh_l0, False, mode)]
    if m.bidirectional:
        outs.append(lstm_dir(x, m.weight_ih_l0_reverse, m.weight_hh_l0_reverse, m.bias_ih_l0_reverse+m.bias_hh_l0_reverse, True, mode))
    return torch.cat(outs,-1)

def block(blk, x, nb_skip, fb_skip, mode, act):
    nb,nt,nf,nc = x.shape
    nb_skip_new_in = x.permute(0,2,1,3).reshape(nb*nf,nt,-1)
    x = x.reshape(nb*nt,nf,-1)
    if not blk.is_first: x = q(x + fb_skip, act)
    x = q(lstm(x, blk.fullLstm, mode), act)
    fb = x
    x = x.view(nb,nt,nf,-1).permute(0,2,1,3).reshape(nb*nf,nt,-1)
    if blk.is_first: x = torch.cat((x, nb_skip_new_in),-1)
    else: x = q(x + nb_skip_new_in, act)
    x = q(lstm(x, blk.narrLstm, mode), act)
    nbs = x
    x = x.view(nb,nf,nt,-1).permute(0,2,1,3)
    return x, fb, nbs

def run(net, x, mode, act):
    x = q(x.permute(0,3,2,1), act)
    nb,nt,nf,nc = x.shape
    x,fb,nbs = block(net.block_1,x,None,None,mode,act)
    x,fb,nbs = block(net.block_2,x,nbs,fb,mode,act)
    x,fb,nbs = block(net.block_3,x,nbs,fb,mode,act)
    x = x.permute(0,2,1,3).reshape(nb*nf,nt,-1)
    ipd = net.pooling(x); ipd = torch.tanh(net.emb2ipd(ipd))
    nt2 = ipd.shape[1]
    ipd = ipd.view(nb,nf,nt2,-1).permute(0,2,1,3)
    return torch.cat((ipd[...,0],ipd[...,1]),2)

if __name__=='__main__':
    torch.set_num_threads(8)
    T = int(sys.argv[1]) if len(sys.argv)>1 else 60
    for online in (True, False):
        torch.manual_seed(0)
        net = Model.FN_SSL(is_online=online).eval()
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(1,4,256,T,generator=g)
        with torch.no_grad():
            ref = net(x)
            for mode,act in (('fp32','fp32'),('fp16','fp32'),('fp16','fp16'),('bf16','fp32')):
                t=time.time(); y = run(net,x,mode,act)
                print(f"online={online} mode={mode} act={act}: maxabs={float((y-ref).abs().max()):.3e} refmax={float(ref.abs().max()):.3e} rel={float((y-ref).abs().max()/ref.abs().max()):.3e} ({time.time()-t:.1f}s)", flush=True)
