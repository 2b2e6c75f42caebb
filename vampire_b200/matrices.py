"""Host-side preparation of the per-camera 4x4 matrices.

Index parity with the reference needs the tiny matrix algebra (``inverse``, ``K @ E^-1``,
``E @ K^-1``) done with the very same torch calls the reference makes
(BV2:334, 340, 372-374, 379), on the device the ``mats_dict`` tensors live on.  The kernels then
receive five row-major 4x4 matrices per (sample, camera):

    slot 0  bda^-1          get_pixel     BV2:372-374   (identity + flag when bda is None)
    slot 1  K @ E^-1        get_pixel     BV2:379
    slot 2  ida             get_pixel     BV2:386-387
    slot 3  ida^-1          get_geometry  BV2:333-334
    slot 4  E @ K^-1        get_geometry  BV2:340
    slot 5  bda             get_geometry  BV2:343-346

i.e. a (B, N, 6, 4, 4) fp32 tensor = 384 B per camera.
"""
from __future__ import annotations

from typing import Optional

import torch

NUM_SLOTS = 6


def prepare_matrices(sensor2ego: torch.Tensor, intrin: torch.Tensor, ida: torch.Tensor,
                     bda: Optional[torch.Tensor]) -> torch.Tensor:
    """(B,N,4,4) x3 [+ (B,4,4)] -> (B,N,6,4,4) fp32 on the inputs' device."""
    B, N = sensor2ego.shape[:2]
    sensor2ego = sensor2ego.float()
    intrin = intrin.float()
    ida = ida.float()
    out = torch.empty(B, N, NUM_SLOTS, 4, 4, dtype=torch.float32, device=sensor2ego.device)
    if bda is not None:
        bda_rep = bda.float().unsqueeze(1).repeat(1, N, 1, 1)
        out[:, :, 0] = bda_rep.view(B, N, 1, 1, 1, 4, 4).inverse().view(B, N, 4, 4)
        out[:, :, 5] = bda_rep
    else:
        eye = torch.eye(4, dtype=torch.float32, device=sensor2ego.device)
        out[:, :, 0] = eye
        out[:, :, 5] = eye
    out[:, :, 1] = intrin.matmul(torch.inverse(sensor2ego))
    out[:, :, 2] = ida
    out[:, :, 3] = ida.view(B, N, 1, 1, 1, 4, 4).inverse().view(B, N, 4, 4)
    out[:, :, 4] = sensor2ego.matmul(torch.inverse(intrin))
    return out.contiguous()
