"""Shared constants of the golden fixtures (importable without the reference tree)."""
TINY = dict(layers=7, width=128, patch=16, grid0=4, embed=64)


def checksum(*tensors):
    return float(sum(t.double().abs().sum().item() for t in tensors))
