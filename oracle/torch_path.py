"""ORACLE (test infrastructure only): torch-CPU functional restatement of the reference path.

Every function issues the same ATen calls, in the same order and on the same shapes, as the
reference method it cites, so on a given host it is bit-identical to the reference.  ``conf`` is
a dict with the reference's ``backbone_conf`` keys (base_exp.py:40-92).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# T1-T3: lattice buffers
# --------------------------------------------------------------------------------------------
def frustum_buffer(conf):
    """BV2:253-271 ``create_frustum`` -> (D, fH, fW, 4) = [u, v, d, 1]."""
    H, W = conf["final_dim"]
    fH, fW = H // conf["downsample_factor"], W // conf["downsample_factor"]
    d = torch.arange(*conf["d_bound"], dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = d.shape[0]
    u = torch.linspace(0, W - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    v = torch.linspace(0, H - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((u, v, d, torch.ones_like(d)), -1)


def voxel_buffer(xb, yb, zb):
    """BV2:273-293 ``create_voxel_coords`` (norm=False) -> (Z, Y, X, 4) = [x, y, z, 1]."""
    def axis(b):
        return torch.linspace(b[0] + b[2] / 2., b[1] - b[2] / 2., int((b[1] - b[0]) / b[2]), dtype=torch.float)
    zs, ys, xs = torch.meshgrid(axis(zb), axis(yb), axis(xb), indexing="ij")
    return torch.stack([xs, ys, zs, torch.ones_like(xs)], dim=-1)


def camera_mids(conf):
    """BV2:243-246."""
    t = torch.arange(*conf["d_bound"], dtype=torch.float)
    return 0.5 * (t[..., :-1] + t[..., 1:])


def bev_mids(conf):
    """BV2:248-251 (flipped: top level first)."""
    zb = conf["z_bound_det"]
    t = torch.linspace(zb[0] + zb[2] / 2., zb[1] - zb[2] / 2., int((zb[1] - zb[0]) / zb[2]), dtype=torch.float)
    return torch.flip(t, dims=[0])


def build_buffers(conf):
    return {
        "frustum": frustum_buffer(conf),
        "voxel_coords": voxel_buffer(conf["x_bound_seg"], conf["y_bound_seg"], conf["z_bound_seg"]),
        "output_coords": voxel_buffer(conf["x_bound_det"], conf["y_bound_det"], conf["z_bound_det"]),
        "camera_mids": camera_mids(conf),
        "bev_mids": bev_mids(conf),
    }


# --------------------------------------------------------------------------------------------
# G1 / G2: projective geometry
# --------------------------------------------------------------------------------------------
def get_geometry(buf, sensor2ego, intrin, ida, bda):
    """BV2:314-349: frustum lattice -> ego xyz, (B, N, D, fH, fW, 3)."""
    B, N = sensor2ego.shape[:2]
    pts = buf["frustum"]
    pts = ida.view(B, N, 1, 1, 1, 4, 4).inverse().matmul(pts.unsqueeze(-1))
    pts = torch.cat((pts[:, :, :, :, :, :2] * pts[:, :, :, :, :, 2:3], pts[:, :, :, :, :, 2:]), 5)
    comb = sensor2ego.matmul(torch.inverse(intrin))
    pts = comb.view(B, N, 1, 1, 1, 4, 4).matmul(pts)
    if bda is not None:
        bda = bda.unsqueeze(1).repeat(1, N, 1, 1).view(B, N, 1, 1, 1, 4, 4)
        pts = (bda @ pts).squeeze(-1)
    else:
        pts = pts.squeeze(-1)
    return pts[..., :3]


def get_pixel(buf, sensor2ego, intrin, ida, bda):
    """BV2:351-388: voxel centres -> augmented-image pixel + camera depth, (B, N, Z, Y, X, 3)."""
    B, N = sensor2ego.shape[:2]
    pts = buf["voxel_coords"]
    if bda is not None:
        bda = bda.unsqueeze(1).repeat(1, N, 1, 1).view(B, N, 1, 1, 1, 4, 4)
        pts = bda.inverse().matmul(pts.unsqueeze(-1))
    else:
        pts = pts.unsqueeze(-1)
    comb = intrin.matmul(torch.inverse(sensor2ego))
    pts = comb.view(B, N, 1, 1, 1, 4, 4).matmul(pts)
    zc = pts[:, :, :, :, :, 2:3]
    pts = torch.cat((pts[..., :2, :] / torch.clamp(zc, min=1e-6), pts[..., 2:, :]), dim=5)
    pts = ida.view(B, N, 1, 1, 1, 4, 4).matmul(pts).squeeze(-1)
    return pts[..., :3]


# --------------------------------------------------------------------------------------------
# L1-L4: lift + pool
# --------------------------------------------------------------------------------------------
def depth_softmax(logits):
    """BV2:551 ``.softmax(dim=1)`` on the (B*N, D, fH, fW) depth logits (autocast runs it in fp32)."""
    return torch.softmax(logits.float(), dim=-3)


def lift_outer(depth, ctx):
    """BV2:553: (B,N,D,h,w) x (B,N,C,h,w) -> (B,N,C,D,h,w)."""
    return depth.unsqueeze(2) * ctx.unsqueeze(3)


def lift_norm_coords(conf, pix):
    """BV2:493-506: validity mask and normalised (clamped) sampling coordinates."""
    H, W = conf["final_dim"]
    db = conf["d_bound"]
    x, y, z = pix[..., 0], pix[..., 1], pix[..., 2]
    xv = (x > -0.5).bool() & (x < float(W - 0.5)).bool()
    yv = (y > -0.5).bool() & (y < float(H - 0.5)).bool()
    zv = (z > db[0]).bool() & (z < db[1]).bool()
    valid = (xv & yv & zv).float()
    nx = 2.0 * (x / float(W - 1)) - 1.0
    ny = 2.0 * (y / float(H - 1)) - 1.0
    nz = 2.0 * ((z - db[0]) / (db[1] - db[0])) - 1.0
    nx = torch.clamp(nx, min=-2.0, max=2.0)
    ny = torch.clamp(ny, min=-2.0, max=2.0)
    nz = torch.clamp(nz, min=-2.0, max=2.0)
    return valid, torch.stack([nx, ny, nz], dim=-1)


def get_voxel_feats(conf, buf, frustum_feats, mats):
    """BV2:483-516: gather the frustum volume at every voxel centre, mean over seeing cameras."""
    B, N, C, d, h, w = frustum_feats.shape
    pix = get_pixel(buf, mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0],
                    mats["ida_mats"][:, 0], mats.get("bda_mat", None))
    Z, Y, X = buf["voxel_coords"].shape[:3]
    valid, nxyz = lift_norm_coords(conf, pix)
    nxyz = nxyz.reshape(-1, Z, Y, X, 3)
    vf = F.grid_sample(frustum_feats.reshape(-1, C, d, h, w), nxyz, align_corners=False)
    vf = vf.reshape(B, N, C, Z, Y, X) * valid.unsqueeze(2)
    mask = (torch.abs(vf) > 0).float()
    numer = torch.sum(vf, dim=1)
    denom = torch.sum(mask, dim=1) + 1e-6
    return numer / denom


def get_voxel_feats_2d(conf, buf, img_feats, mats):
    """base_bilinear.py:471-517 ``BaseBiLinear.get_voxel_feats``: bilinear sampling of the (B,N,C,h,w) image
    features on a depth-1 volume at z = 0, z_valid = z > 0, non-zero mean over cameras."""
    B, N, C, h, w = img_feats.shape
    H, W = conf["final_dim"]
    pix = get_pixel(buf, mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0],
                    mats["ida_mats"][:, 0], mats.get("bda_mat", None))
    Z, Y, X = buf["voxel_coords"].shape[:3]
    x, y, z = pix[..., 0], pix[..., 1], pix[..., 2]
    xv = (x > -0.5).bool() & (x < float(W - 0.5)).bool()
    yv = (y > -0.5).bool() & (y < float(H - 0.5)).bool()
    zv = (z > 0.).bool()
    valid = (xv & yv & zv).float()
    nx = torch.clamp(2.0 * (x / float(W - 1)) - 1.0, min=-2.0, max=2.0)
    ny = torch.clamp(2.0 * (y / float(H - 1)) - 1.0, min=-2.0, max=2.0)
    nxyz = torch.stack([nx, ny, torch.zeros_like(nx)], dim=-1).reshape(-1, Z, Y, X, 3)
    vf = F.grid_sample(img_feats.reshape(-1, C, 1, h, w), nxyz, align_corners=False)
    vf = vf.reshape(B, N, C, Z, Y, X) * valid.unsqueeze(2)
    mask = (torch.abs(vf) > 0).float()
    return torch.sum(vf, dim=1) / (torch.sum(mask, dim=1) + 1e-6)


def lift_pool(conf, buf, depth, ctx, mats):
    """BV2:553 + 563: the fused operation the product implements."""
    return get_voxel_feats(conf, buf, lift_outer(depth, ctx), mats)


# --------------------------------------------------------------------------------------------
# T4: density
# --------------------------------------------------------------------------------------------
def laplace_density(s, beta_param, bias, beta_min=1e-4):
    """render_utils.py:30-46 ``ModifyLaplaceDensity``."""
    beta = beta_param.abs() + beta_min
    alpha = 1 / beta
    return alpha * (0.5 + 0.5 * (s - bias).sign() * torch.expm1(-(s - bias).abs() / beta))


def density(conf, s, beta_param):
    """``self.density`` (BV2:191-194): nn.Sigmoid() for density_mode='naive', ModifyLaplaceDensity for 'sdf'."""
    if conf.get("density_mode", "sdf") == "naive":
        return torch.sigmoid(s)
    return laplace_density(s, beta_param, conf["sdf_bias"])


# --------------------------------------------------------------------------------------------
# R1-R6: volume rendering
# --------------------------------------------------------------------------------------------
def _seg_lo_ext(conf, device=None):
    xb, yb, zb = conf["x_bound_seg"], conf["y_bound_seg"], conf["z_bound_seg"]
    lo = torch.as_tensor([xb[0], yb[0], zb[0]], device=device)
    ext = torch.as_tensor([xb[1] - xb[0], yb[1] - yb[0], zb[1] - zb[0]], device=device)
    return lo, ext


def buffers_to(buf, device):
    """The lattice buffers on another device (bench.py runs these same ATen calls on the B200 as the same-box GPU
    comparator: the reference's own path there is stock ATen CUDA kernels)."""
    return {k: v.to(device) for k, v in buf.items()}


def render_norm_geom(conf, geom):
    """BV2:397-407: normalised sample coordinates of planes 0..D-2 and the inclusive mask."""
    lo, ext = _seg_lo_ext(conf, geom.device)
    g = (geom[:, :, :-1, :, :] - lo) / ext
    g = g * 2. - 1.
    m = (g[..., 0] >= -1.) & (g[..., 0] <= 1.) & (g[..., 1] >= -1.) & (g[..., 1] <= 1.) & \
        (g[..., 2] >= -1.) & (g[..., 2] <= 1.)
    return g, m


def render_norm_output(conf, buf):
    """BV2:408-417."""
    lo, ext = _seg_lo_ext(conf, buf["output_coords"].device)
    g = (buf["output_coords"][..., :3] - lo) / ext
    return g * 2. - 1.


def volume_rendering(conf, buf, geom, density_feature, semantic_logits, voxel_features, rgb, beta_param):
    """BV2:391-467 ``volume_rendering_from_multiple_views`` (both density modes, cat_seg or not)."""
    B, N, d, h, w, _ = geom.shape
    K = conf["num_classes"]
    vol = torch.cat([density_feature, semantic_logits, rgb, voxel_features], dim=1)
    g, m = render_norm_geom(conf, geom)
    go = render_norm_output(conf, buf)
    go = go[None, ...].expand(B, *go.shape)
    ff = F.grid_sample(vol, g.reshape(B, -1, h, w, 3), align_corners=True)
    ff = ff.reshape(B, -1, N, d - 1, h, w).permute(0, 2, 1, 3, 4, 5) * m.unsqueeze(2)
    ff = torch.nan_to_num(ff)
    f_den = density(conf, ff[:, :, :1, ...], beta_param)
    f_seg = ff[:, :, 1:K + 1, ...]
    f_rgb = ff[:, :, K + 1:K + 4, ...]
    f_delta = torch.norm(geom[:, :, 1:, :, :, :] - geom[:, :, :-1, :, :, :], dim=-1)
    sd = f_den * f_delta.unsqueeze(2)
    alpha = 1 - torch.exp(-sd)
    trans = torch.exp(-torch.cat([torch.zeros_like(sd[:, :, :, :1, :, :]),
                                  torch.cumsum(sd[:, :, :, :-1, :, :], dim=3)], dim=3))
    wts = alpha * trans
    acc = wts.sum(dim=3)
    bg_depth = (1 - acc) * conf["d_bound"][1]
    rgb_p = torch.sum(wts * f_rgb, dim=3)
    seg_p = torch.sum(wts * f_seg, dim=3)
    dep_p = (wts * buf["camera_mids"][None, None, None, :, None, None]).sum(dim=3) + bg_depth

    vf = F.grid_sample(vol, go, align_corners=True)
    vf = torch.flip(vf, dims=[2])
    v_den = density(conf, vf[:, :1, ...], beta_param)
    v_seg = vf[:, 1:K + 1, ...]
    v_rgb = vf[:, K + 1:K + 4, ...]
    v_out = vf[:, K + 4:, ...]
    if conf.get("cat_seg", False):
        v_out = torch.cat((v_out, v_seg), dim=1)
    v_delta = torch.ones_like(v_den) * conf["z_bound_det"][2]
    vsd = v_den * v_delta
    v_alpha = 1 - torch.exp(-vsd)
    v_trans = torch.exp(-torch.cat([torch.zeros_like(vsd[:, :, :1, :, :]),
                                    torch.cumsum(vsd[:, :, :-1, :, :], dim=2)], dim=2))
    v_w = v_alpha * v_trans
    bev_rgb = torch.sum(v_w * v_rgb, dim=2)
    bev_seg = torch.sum(v_w * v_seg, dim=2)
    bev_h = (v_w * buf["bev_mids"][None, None, :, None, None]).sum(dim=2)
    return rgb_p, seg_p, dep_p, bev_rgb, bev_seg, bev_h, v_den, v_out


def render_from_mats(conf, buf, mats, density_feature, semantic_logits, voxel_features, rgb, beta_param):
    """BV2:554-559 + 612-614: geometry -> nan_to_num -> render (the fused operation)."""
    geom = get_geometry(buf, mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0],
                        mats["ida_mats"][:, 0], mats.get("bda_mat", None))
    geom = torch.nan_to_num(geom, -1e3)
    return volume_rendering(conf, buf, geom, density_feature, semantic_logits, voxel_features, rgb, beta_param)


# --------------------------------------------------------------------------------------------
# callers right after the path
# --------------------------------------------------------------------------------------------
def upsample(x, factor):
    """BV2:210, 616-626: nn.UpsamplingBilinear2d(scale_factor) on (B*N, ch, fH, fW)."""
    lead = x.shape[:-3]
    y = torch.nn.UpsamplingBilinear2d(scale_factor=factor)(x.reshape(-1, *x.shape[-3:]))
    return y.reshape(*lead, *y.shape[-3:])


def occ_coords(point_cloud_range=(-40.0, -40.0, -1.0, 40.0, 40.0, 5.4), voxel=(0.4, 0.4, 0.4), dims=(200, 200, 16)):
    """BV2:295-302 ``create_norm_occ_coords(norm=False)`` -> (200,200,16,3)."""
    mask = torch.ones(dims, dtype=torch.bool)
    idx = torch.where(mask)
    c = torch.cat((idx[0][:, None] * voxel[0] + voxel[0] / 2 + point_cloud_range[0],
                   idx[1][:, None] * voxel[1] + voxel[1] / 2 + point_cloud_range[1],
                   idx[2][:, None] * voxel[2] + voxel[2] / 2 + point_cloud_range[2]), dim=1)
    return c.reshape(*dims, 3)


def point_queries(conf, semantic_logits, density_feature, pts, i):
    """BV2:579-596 for sample i: (pts_logits (P,K), pts_sdf (P,))."""
    lo, ext = _seg_lo_ext(conf, pts.device)
    n = (pts - lo) / ext
    n = n[None, None, None, :, :]
    n = n * 2. - 1.
    valid = (n[..., 0] >= -1.) & (n[..., 0] <= 1.) & (n[..., 1] >= -1.) & (n[..., 1] <= 1.) & \
            (n[..., 2] >= -1.) & (n[..., 2] <= 1.)
    logits = F.grid_sample(semantic_logits[[i], ...], n, padding_mode='border', align_corners=True)
    sdf = F.grid_sample(density_feature[[i], ...], n, align_corners=True)
    sdf = sdf.squeeze(1) * valid
    return logits[0, :, 0, 0, :].permute(1, 0), sdf[0, 0, 0, :]


def occupancy_queries(conf, semantic_logits, density_feature, bda, beta_param, coords=None):
    """BV2:597-609 + 647-648: (occ_logits (B,X,Y,Z,K), tanh(occ_density) (B,X,Y,Z,1))."""
    coords = (occ_coords() if coords is None else coords).to(bda.device)
    B = semantic_logits.shape[0]
    lo, ext = _seg_lo_ext(conf, bda.device)
    rot = bda[:, :3, :3].view(B, 1, 1, 1, 3, 3)
    c = (rot @ coords[None, ..., None].expand(B, *coords.shape, 1)).squeeze(-1)
    n = (c - lo) / ext
    n = n * 2. - 1.
    logits = F.grid_sample(semantic_logits, n, padding_mode='border', align_corners=True)
    dens = F.grid_sample(density(conf, density_feature, beta_param), n, align_corners=True)
    return logits.permute(0, 2, 3, 4, 1), dens.permute(0, 2, 3, 4, 1).tanh()


# --------------------------------------------------------------------------------------------
# the caller of the path: BV2:518-649 ``_forward_single_sweep`` restated over a backbone-like object
# --------------------------------------------------------------------------------------------
def forward_single_sweep(bb, sweep_index, sweep_imgs, mats_dict, inrange_pts=None):
    """Call-for-call restatement of ``BaseVAMPIRE2._forward_single_sweep`` (BV2:548-649) on ``bb``: any object with
    the reference backbone's submodules, buffers and path methods (the real backbone; a backbone with
    ``vampire_b200.integration.attach`` applied; the tests' stand-in on the GPU box).  Checked bit for bit against the
    reference's own method in tests/test_oracle_vs_reference.py."""
    batch_size, num_sweeps, num_cams = sweep_imgs.shape[:3]
    img_feats = bb.get_cam_feats(sweep_imgs)
    h, w = img_feats.shape[-2], img_feats.shape[-1]
    source_features = img_feats[:, 0, ...].reshape(batch_size * num_cams, -1, h, w)
    depth = bb.mapping_along_depth(source_features).softmax(dim=1).reshape(batch_size, num_cams, -1, h, w)
    ctx = bb.channel_lower(source_features).reshape(batch_size, num_cams, -1, h, w)
    img_feats_with_depth = depth.unsqueeze(2) * ctx.unsqueeze(3)
    geom_xyz = bb.get_geometry(mats_dict['sensor2ego_mats'][:, sweep_index, ...],
                               mats_dict['intrin_mats'][:, sweep_index, ...],
                               mats_dict['ida_mats'][:, sweep_index, ...], mats_dict.get('bda_mat', None))
    voxel_features = bb.get_voxel_feats(img_feats_with_depth, sweep_index, mats_dict)
    if bb.cat_pos:
        nvc = bb.norm_voxel_coords.permute(3, 0, 1, 2)[None, ...].repeat(batch_size, 1, 1, 1, 1)
        voxel_features = torch.cat([voxel_features, nvc], dim=1)
    base_features = bb.base_conv(voxel_features)
    density_feature = bb.density_conv(base_features)
    semantic_logits = bb.seg_conv(base_features)
    rgb = bb.rgb_conv(base_features)
    dev = bb.camera_mids.device
    lo = torch.as_tensor([bb.x_bound_seg[0], bb.y_bound_seg[0], bb.z_bound_seg[0]], device=dev)
    ext = torch.as_tensor([bb.x_bound_seg[1] - bb.x_bound_seg[0], bb.y_bound_seg[1] - bb.y_bound_seg[0],
                           bb.z_bound_seg[1] - bb.z_bound_seg[0]], device=dev)
    pts_logits_batch, pts_sdf_batch = [], []
    if inrange_pts is not None:
        for i in range(batch_size):
            n = (inrange_pts[i] - lo) / ext
            n = n[None, None, None, :, :]
            n = n * 2. - 1.
            valid = (n[..., 0] >= -1.) & (n[..., 0] <= 1.) & (n[..., 1] >= -1.) & (n[..., 1] <= 1.) & \
                    (n[..., 2] >= -1.) & (n[..., 2] <= 1.)
            pl = F.grid_sample(semantic_logits[[i], ...], n, padding_mode='border', align_corners=True)
            pts_logits_batch.append(pl[0, :, 0, 0, :].permute(1, 0))
            if bb.density_mode == 'sdf':
                ps = F.grid_sample(density_feature[[i], ...], n, align_corners=True)
                ps = ps.squeeze(1) * valid
                pts_sdf_batch.append(ps[0, 0, 0, :])
    bda = mats_dict.get('bda_mat', None)[:, :3, :3].view(batch_size, 1, 1, 1, 3, 3)
    occ = (bda @ bb.occ_coords[None, ..., None].expand(batch_size, *bb.occ_coords.shape, 1)).squeeze(-1)
    nocc = (occ - lo) / ext
    nocc = nocc * 2. - 1.
    occ_logits = F.grid_sample(semantic_logits, nocc, padding_mode='border', align_corners=True)
    occ_density = F.grid_sample(bb.density(density_feature), nocc, align_corners=True)
    geom_xyz = torch.nan_to_num(geom_xyz, -1e3)
    rgb_p, seg_p, dep_p, bev_rgb, bev_seg, bev_h, bev_density, voxel_output = \
        bb.volume_rendering_from_multiple_views(geom_xyz, density_feature, semantic_logits, base_features, rgb)
    up = bb.upsample_factor

    def up2(x):
        return bb.upsample2d(x.reshape(batch_size * num_cams, -1, bb.fH, bb.fW)).reshape(
            batch_size, num_cams, -1, bb.fH * up, bb.fW * up)

    rgb_p, seg_p, dep_p = up2(rgb_p), up2(seg_p), up2(dep_p)
    voxel_output = voxel_output * bev_density.tanh() if bb.density_mode == 'sdf' else voxel_output * bev_density
    vof = bb.voxel_output(voxel_output.reshape(batch_size, -1, voxel_output.shape[-2], voxel_output.shape[-1])).float()
    return (vof.contiguous(), rgb_p, seg_p, dep_p, bev_rgb, bev_seg, bev_h, bev_density, pts_logits_batch,
            pts_sdf_batch, occ_logits.permute(0, 2, 3, 4, 1), occ_density.permute(0, 2, 3, 4, 1).tanh())
