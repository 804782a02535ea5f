"""TEST-ONLY: bench.py imports torch.distributed unconditionally; the single-rank emulator run never calls it."""
