import sys, os, math, torch
sys.path.insert(0,'3dgan-inversion_b200'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import eg3d_oracle as oracle, b200eg3d as b2
def relerr(a,b):
    a,b=a.detach().cpu().double(),b.detach().cpu().double(); return ((a-b).norm()/b.norm().clamp_min(1e-30)).item()
g=torch.Generator().manual_seed(5)
n,res,npts=2,16,1000
planes=torch.randn(n,3,32,res,res,generator=g); coords=(torch.rand(n,npts,3,generator=g)-0.5)*1.3
dec=b2.OSGDecoder(32,{'decoder_lr_mul':1,'decoder_output_dim':32}); P={}
with torch.no_grad():
    for k,v in dec.named_parameters():
        v.copy_(torch.randn(v.shape,generator=g)*(0.3 if k.endswith('bias') else 1.0)); P['decoder.'+k]=v.detach().clone()
d_rgb=torch.randn(n,npts,32,generator=g); d_sig=torch.randn(n,npts,1,generator=g); rk={'box_warp':1.2}
Pr={k:v.clone().requires_grad_(True) for k,v in P.items()}
pr,cr=planes.clone().requires_grad_(True),coords.clone().requires_grad_(True)
rgb_ref,sig_ref=oracle.run_model(Pr,pr,cr,rk); (rgb_ref*d_rgb).sum().add((sig_ref*d_sig).sum()).backward()
dec=dec.cuda()
for p in dec.parameters(): p.requires_grad_(True)
pl=planes.permute(0,3,4,1,2).reshape(n,res,res,96).contiguous().cuda().requires_grad_(True); cc=coords.cuda().requires_grad_(True)
out=b2.ImportanceRenderer().run_model(pl,dec,cc,None,rk)
(out['rgb']*d_rgb.cuda()).sum().add((out['sigma']*d_sig.cuda()).sum()).backward()
print('rgb', (out['rgb'].cpu()-rgb_ref).abs().max().item(), 'sigma', (out['sigma'].cpu()-sig_ref).abs().max().item())
print('dplanes', relerr(pl.grad, pr.grad.permute(0,3,4,1,2).reshape(n,res,res,96)), 'dcoords', relerr(cc.grad, cr.grad))
for k,v in dec.named_parameters(): print(k, relerr(v.grad, Pr['decoder.'+k].grad))
