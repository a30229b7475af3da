"""Shim of `denoising_diffusion_pytorch.attend.Attend` for flash=False, dropout=0, scale=None.

PARITY UNPINNED: the real class lives in a pip dependency that is not under /root/reference and
the reference has no tests pinning it.  Published algorithm (non-flash branch): on q,k,v of
shape [b, h, n, d]:  sim = einsum('b h i d, b h j d -> b h i j', q, k) * d**-0.5 ;
attn = softmax(sim, dim=-1) ; out = einsum('b h i j, b h j d -> b h i d', attn, v).
Call site: /root/reference/model.py:352 (constructed at model.py:339).
"""
import torch
import torch.nn as nn


class Attend(nn.Module):
    def __init__(self, dropout=0., flash=False, scale=None):
        super().__init__()
        assert not flash and dropout == 0.
        self.scale = scale

    def forward(self, q, k, v):
        scale = self.scale if self.scale is not None else q.shape[-1] ** -0.5
        sim = torch.einsum('b h i d, b h j d -> b h i j', q, k) * scale
        attn = sim.softmax(dim=-1)
        return torch.einsum('b h i j, b h j d -> b h i d', attn, v)
