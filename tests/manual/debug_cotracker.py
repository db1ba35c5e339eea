"""GPU bring-up aid for csrc/cotracker.cu: compares the workspace intermediates of one CoTracker window with the oracle's.
Test infrastructure (imports oracle/); run on the GPU box:  python tests/manual/debug_cotracker.py"""
import os
import sys
from ctypes import c_int

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sam-pt_b200"))
import torch
import torch.nn.functional as F

from oracle import cotracker_ref as R, pips_ref
from sampt_b200 import native, synth
from sam_pt.point_tracker.cotracker.cotracker import CoTracker, cotracker_shapes

torch.manual_seed(0)
sd = synth.condition_cotracker(synth.make_state_dict(cotracker_shapes(), seed=31))
T, H, W, S = 8, 96, 128, 8
frames = synth.make_clip(T, H, W, seed=5)["frames"].float()
q = torch.tensor([[0.0, 0.30 * W, 0.40 * H], [0.0, 0.72 * W, 0.25 * H], [0.0, 0.55 * W, 0.66 * H], [0.0, 0.18 * W, 0.80 * H],
                  [0.0, 0.97 * W, 0.52 * H]])
N = q.shape[0]
M = N * S
model = CoTracker().cuda()
model.load_state_dict(sd)
model.fnet_on_tensor_cores = False
pyr = model.encode_frames(frames.cuda())
ctx = model.native_context()
lib = native.lib()
fm_o = pyr[0].permute(0, 3, 1, 2).cpu()[None]          # (1,T,128,H4,W4): same features for both sides
H4, W4 = fm_o.shape[-2:]


def al(x):
    return (x + 255) & ~255


def ws_views():
    off = 0
    out = {}
    for name, n in (("xin", M * 456), ("pos", N * 456), ("x", M * 384), ("h", M * 384), ("qkv", M * 1152), ("att", M * 384),
                    ("mlp", M * 1536), ("delta", M * 130)):
        off = al(off)
        out[name] = ctx._ws[off:off + n * 4].view(torch.float32).clone().cpu()
        off += n * 4
    return out


def run(td, sdp, iters, coords0, ff0, tm, vi):
    coords = coords0.clone().cuda()
    ff = ff0.clone().cuda()
    vis = torch.empty((N, S), device="cuda")
    tm_d, vi_d = tm.cuda(), vi.cuda()   # keep alive across the call (temporaries would alias in the caching allocator)
    fidx = torch.tensor([0, 0] + list(range(S)), dtype=torch.int32, device="cuda")
    native.check(lib.sampt_cotracker_window(ctx.handle, native.ptr(pyr[0]), native.ptr(pyr[1]), native.ptr(pyr[2]), native.ptr(pyr[3]),
                                            c_int(H4), c_int(W4), native.ptr(fidx), native.ptr(coords), native.ptr(ff),
                                            native.ptr(tm_d), native.ptr(vi_d), native.ptr(model._time_emb), c_int(N),
                                            c_int(iters), c_int(td), c_int(sdp), native.ptr(vis), native.stream_ptr()), "window")
    torch.cuda.synchronize()
    return coords.cpu(), ff.cpu(), vis.cpu()


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-12)).item()


coords0 = (q[:, 1:] / 4.0)[:, None, :].repeat(1, S, 1)              # (N,S,2)
# perturb so that flows are non-zero and the corr windows differ per slot
coords0 = coords0 + torch.randn(N, S, 2) * 1.5
feat0 = torch.stack([pips_ref.bilinear_sample2d(fm_o[:, 0], coords0[None, n:n + 1, 0, 0], coords0[None, n:n + 1, 0, 1])[0, :, 0]
                     for n in range(N)])                               # (N,128)
ff0 = feat0[:, None, :].repeat(1, S, 1)
tm = (torch.rand(N, S) > 0.3).float()
vi = torch.randn(N, S) * 3

# ---- oracle first-iteration input
coords_o = coords0.permute(1, 0, 2)[None]                              # (1,S,N,2)
ff_o = ff0.permute(1, 0, 2)[None]
pyr_o = pips_ref.build_pyramid(fm_o)
fc = pips_ref.corr_lookup(pyr_o, ff_o, coords_o).permute(0, 2, 1, 3).reshape(N, S, -1)
flows = (coords_o - coords_o[:, 0:1]).permute(0, 2, 1, 3).reshape(N, S, 2)
femb = R.get_2d_embedding(flows, 64)
pos = R.sample_pos_embed((H4, W4), 456, coords_o[:, 0]).reshape(N, 1, 456)
tim = R.time_embed(456, S)[None]
cc = torch.stack([tm, vi], dim=2)
x_in = torch.cat([femb, fc, ff0, cc], dim=2) + pos + tim               # (N,S,456)

c1, f1, v1 = run(1, 1, 1, coords0, ff0, tm, vi)
wsv = ws_views()
xin_g = wsv["xin"].reshape(N, S, 456)
raw = torch.cat([femb, fc, ff0, cc], dim=2)
print("pos      rel err", rel(wsv["pos"].reshape(N, 456), pos[:, 0]))
d = xin_g - x_in
for name, a, b in (("flow emb", 0, 130), ("corr", 130, 326), ("feat", 326, 454), ("mask/vis", 454, 456)):
    print(f"xin[{name}] max abs err", d[:, :, a:b].abs().max().item(), " (values up to", x_in[:, :, a:b].abs().max().item(), ")")

# ---- depth (1,1) internals
p = "updateformer."
x0 = F.linear(x_in, sd[p + "input_transform.weight"], sd[p + "input_transform.bias"])     # (N,S,384)
x1 = R._attn_block(sd, p + "time_blocks.0.", x0)                                          # groups = points
xs = x1.permute(1, 0, 2)                                                                    # (S,N,384) groups = slots
hs = F.layer_norm(xs, (384,), None, None, 1e-6)
qkv_s = F.linear(hs, sd[p + "space_blocks.0.attn.qkv.weight"], sd[p + "space_blocks.0.attn.qkv.bias"])   # (S,N,1152)
x2 = R._attn_block(sd, p + "space_blocks.0.", xs).permute(1, 0, 2)                          # (N,S,384)
qq = qkv_s.reshape(S, N, 3, 8, 48).permute(2, 0, 3, 1, 4)
a = ((qq[0] @ qq[1].transpose(-2, -1)) * 48 ** -0.5).softmax(-1)
att_s = (a @ qq[2]).transpose(1, 2).reshape(S, N, 384).permute(1, 0, 2)                      # (N,S,384)
print("space qkv rel err", rel(wsv["qkv"].reshape(N, S, 1152), qkv_s.permute(1, 0, 2)), "(covers input GEMM, time block, LN, qkv GEMM)")
print("space att rel err", rel(wsv["att"].reshape(N, S, 384), att_s))
print("final x   rel err", rel(wsv["x"].reshape(N, S, 384), x2))
delta = F.linear(x2, sd[p + "flow_head.weight"], sd[p + "flow_head.bias"])
print("delta     rel err", rel(wsv["delta"].reshape(N, S, 130), delta))

# time block internals: depth (2,1) leaves time_blocks.1's qkv / attention output in the workspace
run(2, 1, 1, coords0, ff0, tm, vi)
wsv2 = ws_views()
x2t = x2                                                                                     # input of time_blocks.1
ht1 = F.layer_norm(x2t, (384,), None, None, 1e-6)
qkv_t = F.linear(ht1, sd[p + "time_blocks.1.attn.qkv.weight"], sd[p + "time_blocks.1.attn.qkv.bias"])     # (N,S,1152)
qt = qkv_t.reshape(N, S, 3, 8, 48).permute(2, 0, 3, 1, 4)
at = ((qt[0] @ qt[1].transpose(-2, -1)) * 48 ** -0.5).softmax(-1)
att_t = (at @ qt[2]).transpose(1, 2).reshape(N, S, 384)
print("time qkv  rel err", rel(wsv2["qkv"].reshape(N, S, 1152), qkv_t))
print("time att  rel err", rel(wsv2["att"].reshape(N, S, 384), att_t))
x3 = R._attn_block(sd, p + "time_blocks.1.", x2t)
print("time x    rel err", rel(wsv2["x"].reshape(N, S, 384), x3))

# time block alone: recompute the time block's qkv through a (1,1) run is overwritten; check time attention through x1 instead
ht = F.layer_norm(x0, (384,), None, None, 1e-6)


def oracle_iter(td, sdp, iters):
    orig = R.update_former
    R.update_former = lambda sd_, x, **kw: orig(sd_, x, time_depth=td, space_depth=sdp)
    try:
        preds, vis = R.forward_iteration(sd, fm_o, coords_o, ff_o, vi.t()[None, :, :, None], tm.t()[None, :, :, None] > 0.5, iters)
    finally:
        R.update_former = orig
    return preds[-1][0].permute(1, 0, 2) / 4.0, vis[0].t()


for td, sdp, iters in ((1, 1, 1), (6, 6, 1), (6, 6, 6)):
    cg, fg, vg = run(td, sdp, iters, coords0, ff0, tm, vi)
    co, vo = oracle_iter(td, sdp, iters)
    print(f"depth ({td},{sdp}) iters {iters}: coords max abs err {(cg - co).abs().max().item():.3e} (feature px), vis logits err "
          f"{(vg - vo).abs().max().item():.3e}")
