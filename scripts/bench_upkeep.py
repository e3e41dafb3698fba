#!/usr/bin/env python
"""Time of one density-grid update (NeRFRenderer.update_extra_state) on the GPU: the reference's torch flow (meshgrid /
morton3D / rand / index_put / masked EMA / mean().item() / packbits around the field's density) against the fused path of
csrc/density_grid.cu, for the full sweep (first 16 updates) and the partial update.  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "aaai2023-pvd_b200")]
import torch  # noqa: E402


def main():
    from pvd_b200 import synthetic as syn
    from pvd_b200.fused import HashNeRFField
    torch.manual_seed(0)
    net = HashNeRFField(num_levels=14, desired_resolution=2048).cuda()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    grid, bitfield, _ = syn.lego_bitfield()
    out = {"what": "update_extra_state, hash field, H=128, 1 cascade", "unit": "ms per update"}
    for fused in (False, True):
        for phase, iters in (("full", 0), ("partial", 16)):
            net.density_grid.copy_(torch.from_numpy(grid))
            ts = []
            for rep in range(6):
                net.iter_density = iters
                net.local_step = 0
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                net.update_extra_state(fused=fused)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            out[f"{'fused' if fused else 'torch_flow'}_{phase}"] = round(sorted(ts[1:])[len(ts[1:]) // 2], 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
