"""nmmo_b200 -- B200-native batched Neural MMO simulator behind the reference's vector-env seam."""
from .config import SPEC, ObsLayout, make_config, default_env_args, default_wrapper_args  # noqa: F401

__all__ = ["SPEC", "ObsLayout", "make_config", "default_env_args", "default_wrapper_args"]
