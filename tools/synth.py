"""Synthetic image batches G0/G1 generated directly in HBM with torch.

The generators are defined by a sequential 64-bit LCG; here draw k of an image is obtained in closed form,
s_k = A^k * seed + C * (1 + A + ... + A^(k-1))  (mod 2^64), with the two coefficient tables built once by
doubling, so a 4096-image batch is generated on the GPU in well under a second and is byte-identical to
oracle.generate() (checked by tests/test_gpu_bench_inputs.py).  torch is plumbing here, not the product.
"""
from __future__ import annotations

import torch

A = 6364136223846793005
C = 1442695040888963407


def _wrap(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


class LcgTables:
    """pw[k] = A^k, gs[k] = C * sum_{j<k} A^j as wrapped int64, k = 0..n."""

    def __init__(self, n: int, device):
        pw = torch.ones(1, dtype=torch.int64, device=device)
        gs = torch.zeros(1, dtype=torch.int64, device=device)
        a_m, g_m, m = A, C, 1          # A^m and C * G_m for the current block length m
        while pw.numel() < n + 1:
            # entries m..2m-1 from entries 0..m-1:  A^(m+k) = A^m A^k ;  G_(m+k) = G_m + A^m G_k
            pw = torch.cat([pw, pw * _wrap(a_m)])
            gs = torch.cat([gs, gs * _wrap(a_m) + _wrap(g_m)])
            g_m = (g_m + a_m * g_m) & ((1 << 64) - 1)
            a_m = (a_m * a_m) & ((1 << 64) - 1)
            m *= 2
        self.pw, self.gs = pw[: n + 1].contiguous(), gs[: n + 1].contiguous()

    def draws(self, seed: int, n: int) -> torch.Tensor:
        """First n values of lcg() for `seed` (uint32 range, returned as int64)."""
        s = self.pw[1: n + 1] * _wrap(seed) + self.gs[1: n + 1]
        return (s >> 33) & 0x7FFFFFFF


class Generator:
    def __init__(self, width: int, height: int, device):
        self.W, self.H, self.device = width, height, device
        self.tab = LcgTables(width * height * 3, device)
        y, x = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
        self.base = torch.stack([x * 255 // width, y * 255 // height, (x + y) * 255 // (width + height)], -1)

    def image(self, kind: int, seed: int, out: torch.Tensor | None = None) -> torch.Tensor:
        W, H = self.W, self.H
        if kind == 0:
            img = (self.tab.draws(seed, W * H * 3) & 255).to(torch.uint8).view(H, W, 3)
        elif kind == 1:
            n = (self.tab.draws(seed, W * H) & 15).view(H, W, 1)
            img = ((self.base + n) & 255).to(torch.uint8)
        else:
            raise ValueError("only G0 and G1 are generated on the device")
        if out is not None:
            out.copy_(img)
            return out
        return img

    def batch(self, n: int, first_seed: int = 12345) -> torch.Tensor:
        """BASELINE config 3: image i alternates G0/G1 with seed first_seed + i."""
        out = torch.empty((n, self.H, self.W, 3), dtype=torch.uint8, device=self.device)
        for i in range(n):
            self.image(i % 2, first_seed + i, out[i])
        return out
