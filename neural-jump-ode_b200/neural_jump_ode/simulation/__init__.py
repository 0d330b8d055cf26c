"""Synthetic path generators producing the packed batch format directly on the device
(recurrences of reference simulation/data_generation.py)."""

from .device_paths import (simulate_paths, sample_observations, sample_observations_ragged, make_packed_batch,
                           make_mixed_ragged_batch, concat_batches)
from .conditional_moments import conditional_moments_packed, get_conditional_moments_at_obs

__all__ = ["simulate_paths", "sample_observations", "sample_observations_ragged", "make_packed_batch",
           "make_mixed_ragged_batch", "concat_batches", "conditional_moments_packed", "get_conditional_moments_at_obs"]
