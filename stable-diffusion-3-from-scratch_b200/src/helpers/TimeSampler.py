"""Logit-normal timestep sampler (reference: src/helpers/TimeSampler.py:5-22): sigmoid(randn*s+m)
drawn on the CPU so CPU-oracle and GPU runs see identical t for a given seed."""
import torch


class TimeSampler:
    def __init__(self, weighted=True, m=0.0, s=1.0):
        self.weighted, self.m, self.s = weighted, m, s

    def __call__(self, n):
        return self.sample(n)

    def sample(self, n):
        if self.weighted:
            return torch.sigmoid(torch.randn(n) * self.s + self.m)
        return torch.rand(n)
