"""ddp-pinocchio_b200: B200-native batched DDP hot path (linearization, backward pass, line-search
rollouts) behind a C ABI (include/ddpb.h).  This Python package is host-side plumbing only:
model tables, synthetic problem generation and a ctypes binding.  The directory name carries a
hyphen, so load it with `ddpb_loader.load()` at the repo root (module name `ddp_pinocchio_b200`)."""
from . import models, synthetic, layout, distributed  # noqa: F401
