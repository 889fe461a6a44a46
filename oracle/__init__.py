"""ORACLE: CPU fp32 restatement of the PowerPaint denoising hot path (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package; the product (powerpaint_b200/) never does.
Pinned against the reference's own files run unmodified: net composition (tests/golden/unet_composition.npz), the three
pipelines' `__call__` (tests/golden/pipeline_*_call.npz), the task-prompt token API (tests/golden/token_api.json).
PARITY UNPINNED for the arithmetic inside diffusers' blocks, VAE and schedulers (see oracle/blocks.py).
"""
