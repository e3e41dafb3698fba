"""Encoder factory with the reference's names and defaults (tools/encoding.py:6-123).

`get_encoder(encoding, ...) -> (module, output_dim)` for "None", "frequency", "sphere_harmonics", "hashgrid",
"tiledgrid".  ("ash" in the reference imports a module that does not exist in its repository.)
"""
import torch
import torch.nn as nn


class FreqEncoder(nn.Module):
    """NeRF positional encoding: [x, sin(f_k x), cos(f_k x)]_k with f_k = 2^k (tools/encoding.py:6-49)."""

    def __init__(self, input_dim, max_freq_log2, N_freqs, log_sampling=True, include_input=True,
                 periodic_fns=(torch.sin, torch.cos)):
        super().__init__()
        self.input_dim = input_dim
        self.include_input = include_input
        self.periodic_fns = periodic_fns
        self.output_dim = (input_dim if include_input else 0) + input_dim * N_freqs * len(periodic_fns)
        if log_sampling:
            bands = 2.0 ** torch.linspace(0.0, max_freq_log2, N_freqs)
        else:
            bands = torch.linspace(2.0 ** 0.0, 2.0 ** max_freq_log2, N_freqs)
        self.freq_bands = bands.numpy().tolist()

    def forward(self, input, **kwargs):
        parts = [input] if self.include_input else []
        for f in self.freq_bands:
            for fn in self.periodic_fns:
                parts.append(fn(input * f))
        return torch.cat(parts, dim=-1)


def get_encoder(encoding, input_dim=3, multires=6, degree=4, num_levels=14, level_dim=2, base_resolution=16,
                log2_hashmap_size=19, desired_resolution=4096, align_corners=False, **kwargs):
    if encoding == "None":
        return (lambda x, **kw: x), input_dim
    if encoding == "frequency":
        enc = FreqEncoder(input_dim=input_dim, max_freq_log2=multires - 1, N_freqs=multires, log_sampling=True)
    elif encoding == "sphere_harmonics":
        from shencoder import SHEncoder
        enc = SHEncoder(input_dim=input_dim, degree=degree)
    elif encoding in ("hashgrid", "tiledgrid"):
        from gridencoder import GridEncoder
        enc = GridEncoder(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                          log2_hashmap_size=log2_hashmap_size, desired_resolution=desired_resolution,
                          gridtype="hash" if encoding == "hashgrid" else "tiled", align_corners=align_corners)
    else:
        raise NotImplementedError(f"unknown encoding {encoding!r}")
    return enc, enc.output_dim
