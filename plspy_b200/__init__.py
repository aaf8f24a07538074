"""plspy_b200 -- B200-native resampling engine behind plspy's PLS(...) API."""
__version__ = "0.1.0"
