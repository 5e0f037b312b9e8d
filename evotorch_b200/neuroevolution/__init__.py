from .policy import Policy, count_parameters, fill_parameters, parameter_vector

__all__ = ["Policy", "count_parameters", "fill_parameters", "parameter_vector"]
