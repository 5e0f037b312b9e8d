from typing import Type, Union

import numpy as np
import torch

from . import cloning, hook, misc, ranking, readonlytensor
from .cloning import Clonable, Serializable, deep_clone
from .hook import Hook
from .misc import (
    device_of,
    dtype_of,
    clip_tensor,
    clone,
    empty_tensor_like,
    is_bool,
    is_bool_vector,
    is_dtype_bool,
    is_dtype_float,
    is_dtype_integer,
    is_dtype_object,
    is_dtype_real,
    is_integer,
    is_integer_vector,
    is_real,
    is_real_vector,
    is_sequence,
    numpy_copy,
    to_numpy_dtype,
    ensure_tensor_length_and_dtype,
    make_empty,
    make_gaussian,
    make_I,
    make_nan,
    make_ones,
    make_randint,
    make_tensor,
    make_uniform,
    make_zeros,
    modify_tensor,
    split_workload,
    stdev_from_radius,
    to_stdev_init,
    to_torch_dtype,
)
from .ranking import rank
from .readonlytensor import ReadOnlyTensor, as_read_only_tensor, read_only_tensor, storage_ptr

DType = Union[str, torch.dtype, np.dtype, Type]  # what `dtype=` arguments accept (tools/misc.py `DType`)
Device = Union[str, torch.device]

__all__ = ["cloning", "Clonable", "Serializable", "deep_clone", "ReadOnlyTensor", "as_read_only_tensor", "read_only_tensor", "storage_ptr", "readonlytensor", "clone", "clip_tensor", "empty_tensor_like", "is_bool", "is_bool_vector", "is_dtype_bool", "is_dtype_float", "is_dtype_integer", "is_dtype_object",
           "is_dtype_real", "is_integer", "is_integer_vector", "is_real", "is_real_vector", "is_sequence", "numpy_copy", "to_numpy_dtype", "DType", "Device", "hook", "Hook", "misc", "ranking", "rank", "modify_tensor", "make_gaussian", "make_uniform", "make_empty", "make_zeros", "make_ones", "make_nan", "make_I", "make_randint", "make_tensor", "split_workload", "stdev_from_radius",
           "to_stdev_init", "to_torch_dtype", "ensure_tensor_length_and_dtype"]
