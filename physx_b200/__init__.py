"""physx_b200 -- B200-native implementation of the PhysX rigid-body step hot path."""
from . import scenes  # noqa: F401
