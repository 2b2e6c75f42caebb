#!/usr/bin/env python
"""North-star (c) "TMA-staged voxel tiles" for the camera march, answered with the workload's own geometry.

For the R50 256x704 val-mode rig (the bench workload) this computes, from the reference's get_geometry (oracle, CPU),
what a shared-memory stage of the packed volume would have to hold: for a ray tile (one warp's 8x4 patch, or a CTA of
four patches = 32x4 pixels) and a slab of consecutive samples, the axis-aligned box of voxels covering every
trilinear corner, against the number of DISTINCT voxels the tile actually touches in that slab.

    python tools/tma_box_analysis.py > profiles/r02_tma_march_analysis.md
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import torch_path as tp
from vampire_b200 import synth
from vampire_b200.config import R50_256x704 as cfg

conf = cfg.backbone_kwargs()
buf = tp.build_buffers(conf)
m = synth.make_mats(cfg, 1, "val")
geom = torch.nan_to_num(tp.get_geometry(buf, m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0],
                                        m["bda_mat"]), -1e3)
g, mask = tp.render_norm_geom(conf, geom)                      # (1,N,S,fH,fW,3), (1,N,S,fH,fW)
size = torch.tensor([cfg.vX, cfg.vY, cfg.vZ], dtype=torch.float32)
idx = torch.floor((g + 1) / 2 * (size - 1)).to(torch.int32)[0]  # base corner (x0,y0,z0)
mask = mask[0].numpy()
idx = idx.numpy()
N, S, fH, fW = mask.shape
REC = 48                                                        # bytes per packed bf16 record

print("# Round 2 -- would TMA / shared-memory staging of voxel tiles help the camera march?  (north star (c))\n")
print("Geometry of the bench workload (R50 256x704, val-mode rig, sample 0), from the reference's own get_geometry. "
      "A stage must hold the axis-aligned box of voxels covering all 8 trilinear corners of every in-volume sample of "
      "a ray tile over a slab of consecutive samples (`cp.async.bulk` / TMA move boxes or rows, not scattered "
      "records); the march then reads each record it needs from shared memory.  `box` = voxels in that box, "
      "`distinct` = voxels the tile really touches, both per (tile, slab); bytes at 48 B per bf16 record.\n")
print("| ray tile | samples per slab | slabs with in-volume samples | median box | p90 box | median distinct | "
      "box / distinct (traffic over-fetch), mean | slabs whose box fits 16 KB | 32 KB | 48 KB |")
print("|---|---|---|---|---|---|---|---|---|---|")
rng = np.random.default_rng(0)
for tname, (th, tw) in (("warp: 8x4 pixels", (4, 8)), ("CTA of 4 warps: 32x4 pixels", (4, 32))):
    for ns in (1, 2, 4):
        boxes, dist = [], []
        tiles = [(n, y, x) for n in range(N) for y in range(0, fH, th) for x in range(0, fW, tw)]
        pick = rng.choice(len(tiles), size=min(600, len(tiles)), replace=False)
        for t in pick:
            n, y, x = tiles[t]
            for s0 in range(0, S, ns):
                mk = mask[n, s0:s0 + ns, y:y + th, x:x + tw]
                if not mk.any():
                    continue
                b = idx[n, s0:s0 + ns, y:y + th, x:x + tw][mk]           # (k, 3) base corners
                lo, hi = b.min(0), b.max(0) + 1                           # far corners are base + 1
                boxes.append(int(np.prod(hi - lo + 1)))
                corners = (b[:, None, :] + np.array([[dx, dy, dz] for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)])[None])
                lin = (corners[..., 2].astype(np.int64) * cfg.vY + corners[..., 1]) * cfg.vX + corners[..., 0]
                dist.append(int(np.unique(lin).size))
        boxes, dist = np.array(boxes), np.array(dist)
        fit = [float((boxes * REC <= k * 1024).mean()) for k in (16, 32, 48)]
        print(f"| {tname} | {ns} | {boxes.size} | {int(np.median(boxes))} ({np.median(boxes) * REC / 1024:.1f} KB) | "
              f"{int(np.percentile(boxes, 90))} ({np.percentile(boxes, 90) * REC / 1024:.1f} KB) | {int(np.median(dist))} | "
              f"{(boxes.sum() / dist.sum()):.2f}x | {fit[0]:.0%} | {fit[1]:.0%} | {fit[2]:.0%} |")
print("""
Reading it against what ncu measured for the shipped march (`profiles/r02_ncu_fwd_b8_bf16.md`, `march_fwd_planned`):

* the march already pulls 36 M sectors (1.16 GB) per launch from L2 -- 2.6 TB/s, ~40 % of what the L2 delivers -- and
  its L1 data pipe runs at 65 % of peak; every variant that added load instructions or spread a voxel's bytes over
  more lines got SLOWER (DESIGN.md section 4: density quads 0.457 -> 0.510 ms, 32 + 16-byte split records with 256-bit
  loads 0.455 -> 0.505 ms, L1 prefetch of the value records 0.453 -> 0.484 ms);
* a box stage moves `box / distinct` times the bytes the gathers need (rays cross the grid obliquely, so the box of a
  slab is mostly voxels no ray of the tile touches) -- on top of an L2 that is already the busiest unit -- and only the
  near-range slabs fit a stage that leaves room for 5 resident blocks per SM (2 stages x <= 20 KB);
* shared-memory reads go through the same L1 data pipe as the global loads they would replace (`LDS` and `LDG` share
  `l1tex__data_pipe_lsu_wavefronts`), so the unit that bounds the kernel sees no relief.

The staged variant that WAS built and measured in this repo is the BEV branch's (regular stencil, boxes = whole rows,
over-fetch 1.0x): `bev_channels_tma_kernel`, bit-identical, 0.335 ms against 0.291 ms for the direct-load kernel
(`VB200_BEV_TMA=1`, DESIGN.md section 4).  For the camera march the numbers above say the box is the wrong shape; what did
pay is staging the march's *geometry* instead (the cached plan: 512 contiguous bytes per warp-step, L2-prefetched).
""")
