"""ORACLE: CPU fp32 restatement of the PowerPaint denoising hot path (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package; the product (powerpaint_b200/) never does.
PARITY UNPINNED for the diffusers-side numerics (see oracle/blocks.py); the task-prompt token API
is pinned against the reference's own powerpaint/utils/utils.py (tests/golden/token_api.json).
"""
