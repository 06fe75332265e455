"""msgwam_b200 -- B200-native hot path of python-msgwam (RK3 ray stepping + flux deposition).

``msgwam_b200.libprop`` is the drop-in for the reference's ``lib/libprop.py``;
``msgwam_b200.ensemble`` holds the device-resident ray store for large runs and multi-GPU sharding.
"""
from . import scenarios  # noqa: F401

__all__ = ["libprop", "scenarios", "ensemble"]
