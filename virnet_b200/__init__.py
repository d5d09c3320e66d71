"""virnet_b200 — B200-native (sm_100a) hot path of VIRNet: network forward/backward + ELBO."""
from .networks.VIRNet import VIRAttResUNet, VIRAttResUNetSR  # noqa: F401

__all__ = ["VIRAttResUNet", "VIRAttResUNetSR"]
