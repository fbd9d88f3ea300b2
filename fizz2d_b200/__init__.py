"""fizz2d_b200 -- B200-native batched stepping engine for the Fizz-2D hot path (World.update())."""
from .scene import Scene, FreePolygon, parse_in, load_in, perturbation_deltas, polygon_constants  # noqa: F401
from .batch import BatchedWorld, GroupSpec, format_world, fp64_peak  # noqa: F401
from . import ga, distributed, scenes, raster  # noqa: F401
from .lib import FizzError, ST_RUNNING, ST_CAPPED, ST_NONFINITE, ST_HIT_OVERFLOW  # noqa: F401

__all__ = ["Scene", "FreePolygon", "parse_in", "load_in", "perturbation_deltas", "polygon_constants",
           "BatchedWorld", "GroupSpec", "format_world", "fp64_peak", "FizzError",
           "ST_RUNNING", "ST_CAPPED", "ST_NONFINITE", "ST_HIT_OVERFLOW"]
