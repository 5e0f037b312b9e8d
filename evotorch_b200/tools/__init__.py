from . import misc, ranking
from .misc import (
    ensure_tensor_length_and_dtype,
    make_gaussian,
    make_uniform,
    modify_tensor,
    split_workload,
    stdev_from_radius,
    to_stdev_init,
    to_torch_dtype,
)
from .ranking import rank

__all__ = ["misc", "ranking", "rank", "modify_tensor", "make_gaussian", "make_uniform", "split_workload", "stdev_from_radius",
           "to_stdev_init", "to_torch_dtype", "ensure_tensor_length_and_dtype"]
