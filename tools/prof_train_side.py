"""ncu target for the training-side kernels (geometry / colour augmentation, fused MultiscaleEpe): every kernel is launched
twice (cold, then warm) at the shapes of tools/train_side_bench.py.
  ncu --set full --clock-control none --import-source on -k regex:"geometry_augment|color_|epe_" -o gpurun_out/prof_train_side \
      python tools/prof_train_side.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import augment, losses  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
N, (H, W), (TH, TW) = 8, (540, 960), (448, 832)
i1 = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, device=dev, generator=g)
i2 = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, device=dev, generator=g)
fl = torch.randn(N, 2, H, W, device=dev, generator=g) * 5
mk = torch.full((N, 1, 1, 1), 255, dtype=torch.uint8, device=dev)
geo = augment.GeometryAugmentation(angle_range=(-17, 17), zoom_range=(0.5, 1 / 0.9), aspect_range=(0.9, 1 / 0.9),
                                   translation_range=0.1, target_shape=(TH, TW), orig_shape=(H, W), batch_size=N,
                                   relative_angle=0.25, relative_scale=(0.96, 1 / 0.96), relative_translation=0.25, seed=1)
col = augment.ColorAugmentation(contrast_range=(-0.4, 0.8), brightness_sigma=0.1, channel_range=(0.8, 1.4), batch_size=N,
                                shape=(TH, TW), noise_range=(0, 0.04), saturation=0.5, hue=0.5, seed=2)
P = geo.params(geo.sample()).to(dev)
Pc = col.params(col.sample()).to(dev)
Hl, Wl = 384, 512
preds = [torch.randn(N, 2, Hl // s, Wl // s, device=dev, generator=g).requires_grad_() for s in losses.SCALES]
flow = torch.randn(N, 2, Hl, Wl, device=dev, generator=g) * 3
mask = (torch.rand(N, 1, Hl, Wl, device=dev, generator=g) > 0.2).float()
for _ in range(2):
    a1, a2, _, _ = augment.geometry_augment(i1, i2, fl, mk, P, (TH, TW))
    augment.color_augment(a1, a2, Pc, noise_sigma=0.03, seed=7)
    losses.multiscale_epe(flow, mask, preds).sum().backward()
torch.cuda.synchronize()
